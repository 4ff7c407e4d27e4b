// spb_kernels.cu -- sm_100a kernels of the sp_ path.
//
// Compile: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -lineinfo (see
// __graft_entry__.build()).  -fmad=false is part of the numerical contract: the reference is
// built without FMA contraction and every product/sum here rounds separately like it does.
//
// Memory picture (DESIGN.md "Data layout"): BVH nodes are 128 B (eight 16-byte loads per visit),
// triangles 48 B (three 16-byte loads), all read through the read-only path (LDG.E.128.CONSTANT);
// the bunny/monkey scenes are < 2 MB and live in L1/L2, the environment map (128 MiB) is the
// only HBM-resident read and is touched once per escaping ray.
#include "spb_kernels.cuh"

namespace spb {

std::atomic<unsigned long long> g_kernelLaunches{0};

__device__ __forceinline__ unsigned warp_sum(unsigned v)
{
    return __reduce_add_sync(0xFFFFFFFFu, v);
}

#define SPB_DISPATCH(FN, cfg, ...)                                                              \
    do {                                                                                        \
        int key = ((cfg).math ? 8 : 0) | ((cfg).envFilter ? 4 : 0) | ((cfg).cull ? 2 : 0) |     \
                  ((cfg).stats ? 1 : 0);                                                        \
        switch (key) {                                                                          \
        case 0: FN<0, 0, false, false>(__VA_ARGS__); break;                                     \
        case 1: FN<0, 0, false, true>(__VA_ARGS__); break;                                      \
        case 2: FN<0, 0, true, false>(__VA_ARGS__); break;                                      \
        case 3: FN<0, 0, true, true>(__VA_ARGS__); break;                                       \
        case 4: FN<0, 1, false, false>(__VA_ARGS__); break;                                     \
        case 5: FN<0, 1, false, true>(__VA_ARGS__); break;                                      \
        case 6: FN<0, 1, true, false>(__VA_ARGS__); break;                                      \
        case 7: FN<0, 1, true, true>(__VA_ARGS__); break;                                       \
        case 8: FN<1, 0, false, false>(__VA_ARGS__); break;                                     \
        case 9: FN<1, 0, false, true>(__VA_ARGS__); break;                                      \
        case 10: FN<1, 0, true, false>(__VA_ARGS__); break;                                     \
        case 11: FN<1, 0, true, true>(__VA_ARGS__); break;                                      \
        case 12: FN<1, 1, false, false>(__VA_ARGS__); break;                                    \
        case 13: FN<1, 1, false, true>(__VA_ARGS__); break;                                     \
        case 14: FN<1, 1, true, false>(__VA_ARGS__); break;                                     \
        default: FN<1, 1, true, true>(__VA_ARGS__); break;                                      \
        }                                                                                       \
    } while (0)


// the integrator kernels are instantiated per (math, envFilter) in spb_render.cu
#define SPB_DECL_RENDER(M, E)                                                                   \
    void launch_render_m##M##e##E(const KernelConfig &, const RenderArgs &, cudaStream_t);       \
    void launch_tiles_m##M##e##E(const KernelConfig &, const TileArgs &, cudaStream_t);
SPB_DECL_RENDER(0, 0)
SPB_DECL_RENDER(0, 1)
SPB_DECL_RENDER(1, 0)
SPB_DECL_RENDER(1, 1)

void launch_render(const KernelConfig &cfg, const RenderArgs &args, cudaStream_t stream)
{
    if (args.x1 <= args.x0 || args.y1 <= args.y0) return;
    g_kernelLaunches++;
    if (cfg.math == 0 && cfg.envFilter == 0) launch_render_m0e0(cfg, args, stream);
    else if (cfg.math == 0) launch_render_m0e1(cfg, args, stream);
    else if (cfg.envFilter == 0) launch_render_m1e0(cfg, args, stream);
    else launch_render_m1e1(cfg, args, stream);
}

void launch_tiles_serial(const KernelConfig &cfg, const TileArgs &args, cudaStream_t stream)
{
    if (args.count == 0) return;
    g_kernelLaunches++;
    if (cfg.math == 0 && cfg.envFilter == 0) launch_tiles_m0e0(cfg, args, stream);
    else if (cfg.math == 0) launch_tiles_m0e1(cfg, args, stream);
    else if (cfg.envFilter == 0) launch_tiles_m1e0(cfg, args, stream);
    else launch_tiles_m1e1(cfg, args, stream);
}

// ---------------------------------------------------------------------------------------------
template <bool CULL, bool STATS>
__global__ void __launch_bounds__(256)
k_primary_hits(DScene scene, DCamera cam, uint32_t sample, uint32_t frame, int32_t *tri,
               int32_t *obj, float *tOut, unsigned long long *counters)
{
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned px = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
    const unsigned py = blockIdx.y * 16 + (warp >> 1) * 4 + (lane >> 3);
    uint32_t stack[SPB_STACK_SIZE];
    float stackT[SPB_STACK_SIZE];
    Counters ctr = {0, 0, 0, 0};
    unsigned rays = 0, hits = 0;
    if (px < cam.width && py < cam.height)
    {
        uint32_t index = px + py * cam.width;
        uint32_t rng = stream_seed(index, sample, frame);
        f3 o, d;
        primary_ray(cam, px, py, rng, o, d);
        Hit h = intersect_scene<CULL>(scene, o, d, stack, stackT, STATS ? &ctr : nullptr);
        rays = 1;
        int32_t t = -1;
        if (h.object >= 0)
        {
            t = (int32_t)f2u(ld4(scene.tris + (size_t)h.slot * 3).w);
            hits = h.t > 0.0f ? 1 : 0;
        }
        if (tri) tri[index] = t;
        if (obj) obj[index] = h.object;
        if (tOut) tOut[index] = h.t;
    }
    unsigned sRays = warp_sum(rays), sHits = warp_sum(hits);
    unsigned n = warp_sum(ctr.nodeVisits), tt = warp_sum(ctr.triangleTests), ob = warp_sum(ctr.objectTests);
    if (lane == 0)
    {
        atomicAdd(&counters[CTR_RAYS], (unsigned long long)sRays);
        atomicAdd(&counters[CTR_HITS], (unsigned long long)sHits);
        atomicAdd(&counters[CTR_MISSES], (unsigned long long)(sRays - sHits));
        if (STATS)
        {
            atomicAdd(&counters[CTR_NODE_VISITS], (unsigned long long)n);
            atomicAdd(&counters[CTR_TRIANGLE_TESTS], (unsigned long long)tt);
            atomicAdd(&counters[CTR_OBJECT_TESTS], (unsigned long long)ob);
        }
    }
}

void launch_primary_hits(const KernelConfig &cfg, const DScene &scene, const DCamera &camera,
                         uint32_t sample, uint32_t frame, int32_t *tri, int32_t *obj, float *t,
                         unsigned long long *counters, cudaStream_t stream)
{
    if (camera.width == 0 || camera.height == 0) return;
    g_kernelLaunches++;
    dim3 grid((camera.width + 15) / 16, (camera.height + 15) / 16);
    if (cfg.cull)
    {
        if (cfg.stats) k_primary_hits<true, true><<<grid, 256, 0, stream>>>(scene, camera, sample, frame, tri, obj, t, counters);
        else k_primary_hits<true, false><<<grid, 256, 0, stream>>>(scene, camera, sample, frame, tri, obj, t, counters);
    }
    else
    {
        if (cfg.stats) k_primary_hits<false, true><<<grid, 256, 0, stream>>>(scene, camera, sample, frame, tri, obj, t, counters);
        else k_primary_hits<false, false><<<grid, 256, 0, stream>>>(scene, camera, sample, frame, tri, obj, t, counters);
    }
}

template <bool CULL, bool STATS>
__global__ void __launch_bounds__(128)
k_intersect_batch(DScene scene, uint32_t count, const float *origins3, const float *dirs3,
                  HitRecord *out, unsigned long long *counters)
{
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t stack[SPB_STACK_SIZE];
    float stackT[SPB_STACK_SIZE];
    Counters ctr = {0, 0, 0, 0};
    unsigned rays = 0, hits = 0;
    if (i < count)
    {
        f3 o = mk3(origins3[i * 3], origins3[i * 3 + 1], origins3[i * 3 + 2]);
        f3 d = mk3(dirs3[i * 3], dirs3[i * 3 + 1], dirs3[i * 3 + 2]);
        Hit h = intersect_scene<CULL>(scene, o, d, stack, stackT, STATS ? &ctr : nullptr);
        rays = 1;
        HitRecord r;
        r.t = -1.0f;
        r.materialId = 0;
        r.nx = r.ny = r.nz = 0.0f;
        r.u = r.v = 0.0f;
        r.triangle = -1;
        r.object = -1;
        if (h.object >= 0)
        {
            Surface sf = resolve_hit(scene, h);
            r.t = h.t;
            r.materialId = sf.material;
            r.nx = sf.normal.x; r.ny = sf.normal.y; r.nz = sf.normal.z;
            r.u = sf.uvx; r.v = sf.uvy;
            r.triangle = (int32_t)sf.triangle;
            r.object = h.object;
            hits = h.t > 0.0f ? 1 : 0;
        }
        out[i] = r;
    }
    unsigned lane = threadIdx.x & 31;
    unsigned sRays = warp_sum(rays), sHits = warp_sum(hits);
    unsigned n = warp_sum(ctr.nodeVisits), tt = warp_sum(ctr.triangleTests), ob = warp_sum(ctr.objectTests);
    if (lane == 0 && counters)
    {
        atomicAdd(&counters[CTR_RAYS], (unsigned long long)sRays);
        atomicAdd(&counters[CTR_HITS], (unsigned long long)sHits);
        atomicAdd(&counters[CTR_MISSES], (unsigned long long)(sRays - sHits));
        atomicAdd(&counters[CTR_NODE_VISITS], (unsigned long long)n);
        atomicAdd(&counters[CTR_TRIANGLE_TESTS], (unsigned long long)tt);
        atomicAdd(&counters[CTR_OBJECT_TESTS], (unsigned long long)ob);
    }
}

void launch_intersect_batch(const KernelConfig &cfg, const DScene &scene, uint32_t count,
                            const float *origins3, const float *dirs3, HitRecord *out,
                            unsigned long long *counters, cudaStream_t stream)
{
    if (count == 0) return;
    g_kernelLaunches++;
    unsigned blocks = (count + 127) / 128;
    if (cfg.cull)
    {
        if (cfg.stats) k_intersect_batch<true, true><<<blocks, 128, 0, stream>>>(scene, count, origins3, dirs3, out, counters);
        else k_intersect_batch<true, false><<<blocks, 128, 0, stream>>>(scene, count, origins3, dirs3, out, counters);
    }
    else
    {
        if (cfg.stats) k_intersect_batch<false, true><<<blocks, 128, 0, stream>>>(scene, count, origins3, dirs3, out, counters);
        else k_intersect_batch<false, false><<<blocks, 128, 0, stream>>>(scene, count, origins3, dirs3, out, counters);
    }
}

// sp_RayIntersectMesh (sp_scene.cpp:127-227): result stays in mesh space; the normal is the
// unnormalised e1 x e2 for flat meshes, the normalised interpolation for smooth ones.
template <bool CULL>
__global__ void k_intersect_mesh(DScene scene, uint32_t smooth, const float *origin3,
                                 const float *dir3, HitRecord *out)
{
    uint32_t stack[SPB_STACK_SIZE];
    float stackT[SPB_STACK_SIZE];
    f3 o = mk3(origin3[0], origin3[1], origin3[2]);
    f3 d = mk3(dir3[0], dir3[1], dir3[2]);
    v4u info = ld4u(scene.objInfo);
    float t, u, v;
    uint32_t slot;
    const float inf = u2f(0x7F800000u);
    if (any_nonfinite_inv(d))
        intersect_mesh<CULL, true>(scene, info.x, o, d, inf, stack, stackT, 0, nullptr, t, slot, u, v);
    else
        intersect_mesh<CULL, false>(scene, info.x, o, d, inf, stack, stackT, 0, nullptr, t, slot, u, v);
    HitRecord r;
    r.t = t;
    r.materialId = 0;
    r.nx = r.ny = r.nz = 0.0f;
    r.u = r.v = 0.0f;
    r.triangle = -1;
    r.object = -1;
    if (t >= 0.0f)
    {
        const v4f *tp = scene.tris + (size_t)slot * 3;
        v4f a = ld4(tp + 0), b = ld4(tp + 1), c = ld4(tp + 2);
        uint32_t tri = f2u(a.w);
        const v4f *sp = scene.shade + ((size_t)info.y + tri) * 4;
        v4f s0 = ld4(sp + 0), s1 = ld4(sp + 1), s2 = ld4(sp + 2), s3 = ld4(sp + 3);
        float w = 1.0f - u - v;
        r.u = s0.w * w + s2.w * u + s3.y * v;
        r.v = s1.w * w + s3.x * u + s3.z * v;
        f3 n;
        if (smooth)
        {
            f3 n0 = mk3(s0.x, s0.y, s0.z), n1 = mk3(s1.x, s1.y, s1.z), n2 = mk3(s2.x, s2.y, s2.z);
            n = normalize3(add3(add3(mul3(n0, w), mul3(n1, u)), mul3(n2, v)));
        }
        else
        {
            f3 pa = mk3(a.x, a.y, a.z);
            n = cross3(sub3(mk3(b.x, b.y, b.z), pa), sub3(mk3(c.x, c.y, c.z), pa));
        }
        r.nx = n.x; r.ny = n.y; r.nz = n.z;
        r.triangle = (int32_t)tri;
        r.object = 0;
    }
    *out = r;
}

void launch_intersect_mesh(const KernelConfig &cfg, const DScene &scene, uint32_t smooth,
                           const float *origin3, const float *dir3, HitRecord *out,
                           cudaStream_t stream)
{
    g_kernelLaunches++;
    if (cfg.cull) k_intersect_mesh<true><<<1, 1, 0, stream>>>(scene, smooth, origin3, dir3, out);
    else k_intersect_mesh<false><<<1, 1, 0, stream>>>(scene, smooth, origin3, dir3, out);
}

__global__ void k_collect_leaves(DScene scene, const float *origin3, const float *dir3,
                                 uint32_t *leaves, uint32_t maxLeaves, uint32_t *countAndError)
{
    uint32_t stack[SPB_STACK_SIZE];
    float stackT[1];
    f3 o = mk3(origin3[0], origin3[1], origin3[2]);
    f3 d = mk3(dir3[0], dir3[1], dir3[2]);
    f3 inv = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    v4u info = ld4u(scene.objInfo);
    uint32_t count = 0, error = 0;
    if (info.x != SPB_REF_EMPTY)
    {
        auto leaf = [&](uint32_t slot, float cull) -> float {
            if (count < maxLeaves)
                leaves[count++] = f2u(ld4(scene.tris + (size_t)slot * 3).w);
            else
                error = 1; // bvh.cpp:296-301
            return cull;
        };
        const float inf = u2f(0x7F800000u);
        traverse<false, true>(scene.nodes, info.x, o, inv, inf, stack, stackT, 0, SPB_STACK_SIZE,
                              nullptr, leaf);
    }
    countAndError[0] = count;
    countAndError[1] = error;
}

void launch_collect_leaves(const DScene &scene, const float *origin3, const float *dir3,
                           uint32_t *leaves, uint32_t maxLeaves, uint32_t *countAndError,
                           cudaStream_t stream)
{
    g_kernelLaunches++;
    k_collect_leaves<<<1, 1, 0, stream>>>(scene, origin3, dir3, leaves, maxLeaves, countAndError);
}

// Progressive accumulation across frames (the reference lists it as not done: main.cpp:75
// "Accumulate pixel samples over multiple frames [ ]"): the running mean of the frames rendered
// so far, accum += (frame - accum) / (n + 1) per component, alpha kept at the frame's.
__global__ void __launch_bounds__(256) k_accumulate_frame(v4f *accum, const v4f *frame, uint32_t count, uint32_t n)
{
    const float denom = (float)(n + 1u);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x)
    {
        v4f f = frame[i];
        if (n == 0)
        {
            accum[i] = f;
            continue;
        }
        v4f a = accum[i];
        a.x = a.x + (f.x - a.x) / denom;
        a.y = a.y + (f.y - a.y) / denom;
        a.z = a.z + (f.z - a.z) / denom;
        a.w = f.w;
        accum[i] = a;
    }
}

void launch_accumulate_frame(v4f *accum, const v4f *frame, uint32_t count, uint32_t framesAccumulated, cudaStream_t stream)
{
    g_kernelLaunches++;
    unsigned grid = (count + 255) / 256;
    if (grid > 148u * 16u) grid = 148u * 16u;
    k_accumulate_frame<<<grid, 256, 0, stream>>>(accum, frame, count, framesAccumulated);
}

// Batched forms for the reference's own performance tests (perf_tests/perf_tests.cpp:51-118 TestBvh:
// bvh_IntersectRay over seeded boxes; :212-305 TestMeshMidphase: sp_RayIntersectMesh over a seeded
// icosphere): one ray per thread, object-space rays against object 0's mesh.
// Leaves: per ray the number of leaves whose own box the ray passes and the XOR and the sum of their
// primitive indices (a fingerprint the checker's leaf list can be compared with).
__global__ void __launch_bounds__(128)
k_collect_leaves_batch(DScene scene, uint32_t count, const float *origins3, const float *dirs3, uint32_t *out3)
{
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= count) return;
    uint32_t stack[SPB_STACK_SIZE];
    float stackT[1];
    f3 o = mk3(origins3[q * 3 + 0], origins3[q * 3 + 1], origins3[q * 3 + 2]);
    f3 d = mk3(dirs3[q * 3 + 0], dirs3[q * 3 + 1], dirs3[q * 3 + 2]);
    f3 inv = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    v4u info = ld4u(scene.objInfo);
    uint32_t n = 0, x = 0, sum = 0;
    if (info.x != SPB_REF_EMPTY)
    {
        auto leaf = [&](uint32_t slot, float cull) -> float {
            uint32_t prim = f2u(ld4(scene.tris + (size_t)slot * 3).w);
            n++;
            x ^= prim * 0x9E3779B1u;
            sum += prim;
            return cull;
        };
        const float inf = u2f(0x7F800000u);
        if (any_nonfinite_inv(d))
            traverse<false, true>(scene.nodes, info.x, o, inv, inf, stack, stackT, 0, SPB_STACK_SIZE, nullptr, leaf);
        else
            traverse<false, false>(scene.nodes, info.x, o, inv, inf, stack, stackT, 0, SPB_STACK_SIZE, nullptr, leaf);
    }
    out3[(size_t)q * 3 + 0] = n;
    out3[(size_t)q * 3 + 1] = x;
    out3[(size_t)q * 3 + 2] = sum;
}

void launch_collect_leaves_batch(const DScene &scene, uint32_t count, const float *origins3, const float *dirs3,
                                 uint32_t *out3, cudaStream_t stream)
{
    g_kernelLaunches++;
    k_collect_leaves_batch<<<(count + 127) / 128, 128, 0, stream>>>(scene, count, origins3, dirs3, out3);
}

// sp_RayIntersectMesh (sp_scene.cpp:127-227) per ray: t (or -1) and the triangle index
template <bool CULL>
__global__ void __launch_bounds__(128)
k_intersect_mesh_batch(DScene scene, uint32_t count, const float *origins3, const float *dirs3, float *tOut, int32_t *triOut)
{
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= count) return;
    uint32_t stack[SPB_STACK_SIZE];
    float stackT[SPB_STACK_SIZE];
    f3 o = mk3(origins3[q * 3 + 0], origins3[q * 3 + 1], origins3[q * 3 + 2]);
    f3 d = mk3(dirs3[q * 3 + 0], dirs3[q * 3 + 1], dirs3[q * 3 + 2]);
    v4u info = ld4u(scene.objInfo);
    float t, u, v;
    uint32_t slot;
    const float inf = u2f(0x7F800000u);
    if (any_nonfinite_inv(d))
        intersect_mesh<CULL, true>(scene, info.x, o, d, inf, stack, stackT, 0, nullptr, t, slot, u, v);
    else
        intersect_mesh<CULL, false>(scene, info.x, o, d, inf, stack, stackT, 0, nullptr, t, slot, u, v);
    tOut[q] = t;
    triOut[q] = t >= 0.0f ? (int32_t)f2u(ld4(scene.tris + (size_t)slot * 3).w) : -1;
}

void launch_intersect_mesh_batch(const KernelConfig &cfg, const DScene &scene, uint32_t count, const float *origins3,
                                 const float *dirs3, float *tOut, int32_t *triOut, cudaStream_t stream)
{
    g_kernelLaunches++;
    const unsigned grid = (count + 127) / 128;
    if (cfg.cull) k_intersect_mesh_batch<true><<<grid, 128, 0, stream>>>(scene, count, origins3, dirs3, tOut, triOut);
    else k_intersect_mesh_batch<false><<<grid, 128, 0, stream>>>(scene, count, origins3, dirs3, tOut, triOut);
}

// simd_RayIntersectAabb4 (simd.h:198-271) on the device, query by query: the three forms of the slab
// test the kernels use, evaluated on the caller's boxes and (origin, reciprocal direction) exactly as
// the reference's function takes them.  masks[q * 3 + 0] = slab_exact (the SSE semantics: the second
// operand of min/max wins on NaN), [1] = slab_fast (hardware min/max; identical when no product is
// NaN), [2] = the conservative test of the resumable traversal (slab_wide, padded by the query's
// own extent) or 0xFFFFFFFF when the ray is one that machine hands to the exact walk (non-finite or
// huge reciprocal).  Bit k = box k.  tnear[q * 4 + k] = entry distance of slab_exact.
__global__ void k_slab_kat(uint32_t count, const float *boxMin12, const float *boxMax12, const float *origin3,
                           const float *invDir3, uint32_t *masks, float *tnear)
{
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= count) return;
    const float *mn = boxMin12 + (size_t)q * 12, *mx = boxMax12 + (size_t)q * 12;
    const f3 o = mk3(origin3[q * 3 + 0], origin3[q * 3 + 1], origin3[q * 3 + 2]);
    const f3 inv = mk3(invDir3[q * 3 + 0], invDir3[q * 3 + 1], invDir3[q * 3 + 2]);
    uint32_t exact = 0, fast = 0, wide = 0;
    float extent = 0.0f;
    for (int k = 0; k < 12; ++k) extent = fmaxf(extent, fmaxf(fabsf(mn[k]), fabsf(mx[k])));
    const f3 d = mk3(1.0f / inv.x, 1.0f / inv.y, 1.0f / inv.z);
    const bool machineRay = trav2_ray_ok(o, d, extent) && fabsf(inv.x) <= 1.0e30f && fabsf(inv.y) <= 1.0e30f && fabsf(inv.z) <= 1.0e30f;
    Trav2 st;
    st.p = st.q = st.cn = st.cf = mk3(0.0f, 0.0f, 0.0f);
    if (machineRay) trav2_constants(o, inv, extent, st);
    for (int k = 0; k < 4; ++k)
    {
        float tn = 0.0f, tf = 0.0f;
        if (slab_exact(mn[k * 3 + 0], mn[k * 3 + 1], mn[k * 3 + 2], mx[k * 3 + 0], mx[k * 3 + 1], mx[k * 3 + 2], o, inv, tn)) exact |= 1u << k;
        tnear[(size_t)q * 4 + k] = tn;
        if (slab_fast(mn[k * 3 + 0], mn[k * 3 + 1], mn[k * 3 + 2], mx[k * 3 + 0], mx[k * 3 + 1], mx[k * 3 + 2], o, inv, tf)) fast |= 1u << k;
        if (machineRay && slab_wide(mn[k * 3 + 0], mn[k * 3 + 1], mn[k * 3 + 2], mx[k * 3 + 0], mx[k * 3 + 1], mx[k * 3 + 2], st,
                                    u2f(0x7F800000u)) < u2f(0x7F800000u))
            wide |= 1u << k;
    }
    masks[(size_t)q * 3 + 0] = exact;
    masks[(size_t)q * 3 + 1] = fast;
    masks[(size_t)q * 3 + 2] = machineRay ? wide : 0xFFFFFFFFu;
}

void launch_slab_kat(uint32_t count, const float *boxMin12, const float *boxMax12, const float *origin3, const float *invDir3,
                     uint32_t *masks, float *tnear, cudaStream_t stream)
{
    g_kernelLaunches++;
    k_slab_kat<<<(count + 127) / 128, 128, 0, stream>>>(count, boxMin12, boxMax12, origin3, invDir3, masks, tnear);
}

template <int MATH, int ENVFILTER, bool CULL, bool STATS>
__global__ void k_radiance_for_path(const DMaterials *materials, const float *path15, uint32_t n,
                                    float clampValue, float *out3)
{
    const DMaterials &M = *materials;
    f3 radiance = mk3(0.0f, 0.0f, 0.0f);
    for (int i = (int)n - 1; i >= 0; --i)
    {
        const float *p = path15 + (size_t)i * 15;
        uint32_t materialId = f2u(p[0]);
        f3 V = mk3(p[4], p[5], p[6]);
        f3 L = mk3(p[7], p[8], p[9]);
        f3 N = mk3(p[10], p[11], p[12]);
        VertexTerms vt = vertex_terms<MATH, ENVFILTER>(M, materialId, L, N, V, p[13], p[14], nullptr);
        radiance = fold_radiance(vt, radiance, clampValue);
    }
    out3[0] = radiance.x;
    out3[1] = radiance.y;
    out3[2] = radiance.z;
}

template <int MATH, int ENVFILTER, bool CULL, bool STATS>
static void launch_radiance_t(const DMaterials *materials, const float *path15, uint32_t n,
                              float clampValue, float *out3, cudaStream_t stream)
{
    k_radiance_for_path<MATH, ENVFILTER, CULL, STATS><<<1, 1, 0, stream>>>(materials, path15, n, clampValue, out3);
}

void launch_radiance_for_path(const KernelConfig &cfg, const DMaterials *materials,
                              const float *path15, uint32_t n, float clampValue, float *out3,
                              cudaStream_t stream)
{
    g_kernelLaunches++;
    KernelConfig c = cfg;
    c.cull = 0;
    c.stats = 0;
    SPB_DISPATCH(launch_radiance_t, c, materials, path15, n, clampValue, out3, stream);
}

template <int MATH, int ENVFILTER, bool CULL, bool STATS>
__global__ void k_evaluate_material(const DMaterials *materials, uint32_t slot,
                                    const float *vertex15, float *out7)
{
    const DMaterials &M = *materials;
    f3 V = mk3(vertex15[4], vertex15[5], vertex15[6]);
    MaterialOut mo = evaluate_material<MATH, ENVFILTER>(M, M.keys[slot], V, vertex15[13], vertex15[14], nullptr);
    out7[0] = mo.albedo.x; out7[1] = mo.albedo.y; out7[2] = mo.albedo.z;
    out7[3] = mo.emission.x; out7[4] = mo.emission.y; out7[5] = mo.emission.z;
    out7[6] = mo.roughness;
}

template <int MATH, int ENVFILTER, bool CULL, bool STATS>
static void launch_evalmat_t(const DMaterials *materials, uint32_t slot, const float *vertex15,
                             float *out7, cudaStream_t stream)
{
    k_evaluate_material<MATH, ENVFILTER, CULL, STATS><<<1, 1, 0, stream>>>(materials, slot, vertex15, out7);
}

void launch_evaluate_material(const KernelConfig &cfg, const DMaterials *materials,
                              uint32_t materialSlot, const float *vertex15, float *out7,
                              cudaStream_t stream)
{
    g_kernelLaunches++;
    KernelConfig c = cfg;
    c.cull = 0;
    c.stats = 0;
    SPB_DISPATCH(launch_evalmat_t, c, materials, materialSlot, vertex15, out7, stream);
}

// ---------------------------------------------------------------------------------------------
// Output stage: the tone mapping the reference applies when it shows the
// path tracer's image (post_processing.frag.glsl:19-26: color *= exposure; color /= 1 + color;
// pow(color, 1/2.2)) and the 8-bit UNORM store of the colour attachment, r in the low byte like
// ToColor (math_lib.h:523-532).  MATH == 0: pow evaluated in double and rounded once (what the
// oracle's deterministic-math build computes); MATH == 1: powf.
template <int MATH>
__global__ void __launch_bounds__(256)
k_tone_map(const v4f *pixels, uint32_t count, float exposure, uint32_t *out)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x)
    {
        v4f p = pixels[i];
        float c[3] = {p.x, p.y, p.z};
        uint32_t packed = 0xFF000000u;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch)
        {
            float v = c[ch] * exposure;
            v = v / (1.0f + v);
            v = MATH == 0 ? (float)pow((double)v, 1.0 / 2.2) : powf(v, 1.0f / 2.2f);
            float q = v > 0.0f ? (v < 1.0f ? v : 1.0f) : 0.0f; // NaN -> 0
            packed |= (uint32_t)floorf(q * 255.0f + 0.5f) << (8 * ch);
        }
        out[i] = packed;
    }
}

void launch_tone_map(const KernelConfig &cfg, const v4f *pixels, uint32_t count, float exposure, uint32_t *out,
                     cudaStream_t stream)
{
    g_kernelLaunches++;
    int device = 0, sms = 0;
    cudaGetDevice(&device);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    unsigned want = (count + 255u) / 256u, grid = (unsigned)sms * 8u;
    if (want < grid) grid = want ? want : 1u;
    if (cfg.math) k_tone_map<1><<<grid, 256, 0, stream>>>(pixels, count, exposure, out);
    else k_tone_map<0><<<grid, 256, 0, stream>>>(pixels, count, exposure, out);
}

} // namespace spb
