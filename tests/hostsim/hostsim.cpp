// tests/hostsim/hostsim.cpp -- DEBUGGING AID, NOT PRODUCT, NOT AN ORACLE.
//
// Compiles the device arithmetic (vk_cinematic_b200/csrc/spb_core.cuh) and the host builder for
// the host with g++, behind the oracle harness ABI (oracle/ora_api.h), so the traversal and
// shading *logic* of the CUDA path can be compared bit-for-bit with the oracle in a container
// that has no GPU.  The shipped library (libspb200.so) contains none of this and has no CPU
// path; tests that use this file say "hostsim" in their name and make no parity claim for the
// GPU -- those are the `-m gpu` tests.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "../../vk_cinematic_b200/csrc/spb_capture.h"
#include "../../vk_cinematic_b200/csrc/spb_cubemap.cuh"
#include "../../oracle/ora_api.h"

using namespace spb;

struct ora_Scene
{
    std::vector<std::shared_ptr<MeshAccel>> meshes;
    std::vector<uint32_t> meshSmooth;
    std::vector<ObjectInstance> objects;
    FlatScene flat;
    DScene d;
    sp_MaterialSystem ms;
    sp_Camera camera;
    ImagePlane plane;
    DMaterials dm;
    DCamera dc;
};

static void refresh(ora_Scene *s)
{
    const v4f *pix[SPB_MAX_IMAGES] = {};
    for (uint32_t i = 0; i < s->ms.imageCount && i < SPB_MAX_IMAGES; ++i)
        pix[i] = (const v4f *)s->ms.images[i].pixels;
    convert_materials(&s->ms, pix, &s->dm);
    convert_camera(&s->camera, &s->dc);
}

extern "C" const char *ora_name(void) { return "hostsim"; }
extern "C" uint32_t ora_max_bounces(void) { return SPB_MAX_BOUNCES; }

extern "C" ora_Scene *ora_create(void)
{
    ora_Scene *s = new ora_Scene();
    memset(&s->ms, 0, sizeof(s->ms));
    memset(&s->camera, 0, sizeof(s->camera));
    memset(&s->plane, 0, sizeof(s->plane));
    s->camera.imagePlane = &s->plane;
    s->flat = flatten_scene(s->objects);
    return s;
}
extern "C" void ora_destroy(ora_Scene *s) { delete s; }

// 1: mesh trees come from the host emulation of the device LBVH builder (spb_lbvh.cuh functions in a
// loop + bvh4_from_binary), 2: the same binary tree collapsed by the host emulation of the DEVICE collapse
// (lbvh_dp / lbvh_emit level by level + bvh4_adopt_device_tree), 0: the host SAH builder
static int g_hostsimBuilder = 0;
extern "C" void hostsim_set_builder(int builder) { g_hostsimBuilder = builder; }
typedef Bvh4 (*BuilderFn)(const float *, const float *, uint32_t);
static BuilderFn builder_fn(int builder)
{
    return builder == 2 ? &build_bvh4_lbvh_host_device_collapse : builder == 1 ? &build_bvh4_lbvh_host : nullptr;
}

// out8: nodeCount, maxDepth, stackNeed, leafCount, fellBack (1 if the LBVH path refused and the SAH
// builder made the tree), summed node half-area x 1e3 (a cost proxy), 0, 0
extern "C" void hostsim_build_info(const float *vertices, uint32_t vertexCount, const uint32_t *indices,
                                   uint32_t indexCount, int builder, uint32_t *out8)
{
    auto accel = build_mesh_accel((const VertexPNT *)vertices, vertexCount, indices, indexCount,
                                  builder_fn(builder));
    const Bvh4 &b = accel->bvh;
    out8[0] = (uint32_t)b.nodes.size();
    out8[1] = b.maxDepth;
    out8[2] = b.stackNeed;
    out8[3] = (uint32_t)b.slotPrim.size();
    out8[4] = 0;
    if (builder)
    {
        auto sah = build_mesh_accel((const VertexPNT *)vertices, vertexCount, indices, indexCount, nullptr);
        out8[4] = sah->bvh.nodes.size() == b.nodes.size() && sah->bvh.maxDepth == b.maxDepth &&
                  memcmp(sah->bvh.nodes.data(), b.nodes.data(), b.nodes.size() * sizeof(Node4)) == 0;
    }
    double area = 0.0;
    for (const Node4 &n : b.nodes)
        for (int k = 0; k < 4; ++k)
        {
            if (n.ref[k] == SPB_REF_EMPTY || (n.ref[k] & SPB_REF_LEAF)) continue;
            double dx = n.bmax[0][k] - n.bmin[0][k], dy = n.bmax[1][k] - n.bmin[1][k], dz = n.bmax[2][k] - n.bmin[2][k];
            area += dx * dy + dy * dz + dz * dx;
        }
    out8[5] = (uint32_t)(area * 1e3);
    out8[6] = out8[7] = 0;
}

// bvh4_adopt_device_tree on a deliberately damaged tree (what a faulty device pass could return): builds the
// emulated device tree of `count` boxes, applies damage `kind` and returns what the check says (1 = accepted).
// kind 0: none; 1: a leaf slot referenced twice; 2: a node referenced twice; 3: a reference to an earlier node
// (a cycle); 4: child count that hides a child; 5: wrong depth; 6: a primitive twice in slotPrim; 7: a child box
// that sticks out of the box its parent holds for the node; 8: a reference beyond the node array; 9: an
// unreferenced (orphan) node appended.
extern "C" int hostsim_adopt_damaged(const float *aabbMin, const float *aabbMax, uint32_t count, int kind)
{
    DeviceTree4 t;
    if (!lbvh_collapse_host_emulation(aabbMin, aabbMax, count, lbvh_build_binary_host(aabbMin, aabbMax, count), &t)) return -1;
    const uint32_t nodes = (uint32_t)(t.nodes.size() / 32);
    auto ref = [&](uint32_t n, uint32_t k) -> uint32_t & { return t.nodes[(size_t)n * 32 + 24 + k]; };
    // first node with an internal child / a leaf child beyond the root
    uint32_t withInner = 0xFFFFFFFFu, innerK = 0, withLeaf = 0xFFFFFFFFu, leafK = 0;
    for (uint32_t n = 0; n < nodes; ++n)
        for (uint32_t k = 0; k < t.nodes[(size_t)n * 32 + 28]; ++k)
        {
            if ((ref(n, k) & SPB_REF_LEAF) && withLeaf == 0xFFFFFFFFu) { withLeaf = n; leafK = k; }
            if (!(ref(n, k) & SPB_REF_LEAF) && withInner == 0xFFFFFFFFu && n > 0) { withInner = n; innerK = k; }
        }
    if (withLeaf == 0xFFFFFFFFu || withInner == 0xFFFFFFFFu) return -2;
    switch (kind)
    {
    case 1: ref(withInner, innerK) = ref(withLeaf, leafK); break;
    case 2: ref(withLeaf, leafK) = ref(withInner, innerK); break;
    case 3: ref(withInner, innerK) = 0; break;
    case 4: t.nodes[(size_t)withInner * 32 + 28] -= 1; break;
    case 5: t.nodes[(size_t)withInner * 32 + 29] += 1; break;
    case 6: t.slotPrim[1] = t.slotPrim[0]; break;
    case 7:
    {
        // the box the parent holds for its child shrinks to nothing: the child's own children stick out of it
        // (a damaged LEAF box would simply be healed: leaf boxes are overwritten with the caller's AABBs)
        float tiny = -3.0e30f;
        memcpy(&t.nodes[(size_t)withInner * 32 + 12 + innerK], &tiny, 4); // bmax[0][innerK]
        break;
    }
    case 8: ref(withInner, innerK) = nodes + 7; break;
    case 9: t.nodes.resize(t.nodes.size() + 32, 0); t.nodes[t.nodes.size() - 32 + 28] = 1; break;
    default: break;
    }
    Bvh4 out;
    return bvh4_adopt_device_tree(aabbMin, aabbMax, count, t, &out) ? 1 : 0;
}

extern "C" int ora_add_mesh(ora_Scene *s, const float *vertices, uint32_t vertexCount,
                            const uint32_t *indices, uint32_t indexCount, uint32_t smooth)
{
    s->meshes.push_back(build_mesh_accel((const VertexPNT *)vertices, vertexCount, indices, indexCount,
                                         builder_fn(g_hostsimBuilder)));
    s->meshSmooth.push_back(smooth);
    return (int)s->meshes.size() - 1;
}

extern "C" int ora_add_object(ora_Scene *s, uint32_t mesh, uint32_t material, const float *p,
                              const float *q, const float *sc)
{
    ObjectInstance ob;
    ob.mesh = s->meshes[mesh];
    ob.material = material;
    ob.smooth = s->meshSmooth[mesh];
    compute_object_transform(ob.mesh.get(), ob.mesh->vertices.data(),
                             (uint32_t)ob.mesh->vertices.size(), vec3{p[0], p[1], p[2]},
                             quat{q[0], q[1], q[2], q[3]}, vec3{sc[0], sc[1], sc[2]}, &ob.model,
                             &ob.invModel, ob.aabbMin, ob.aabbMax);
    s->objects.push_back(ob);
    return (int)s->objects.size() - 1;
}

// 1: the watertight triangle test (sp_b200_Params::triangleTest) in every later ora_build
static uint32_t g_hostsimTriangleTest = 0;
extern "C" void hostsim_set_triangle_test(uint32_t test) { g_hostsimTriangleTest = test; }

extern "C" void ora_build(ora_Scene *s)
{
    s->flat = flatten_scene(s->objects);
    s->d.nodes = s->flat.nodes.data();
    s->d.tris = s->flat.tris.data();
    s->d.shade = s->flat.shade.data();
    s->d.objInv = s->flat.objInv.data();
    s->d.objModel = s->flat.objModel.data();
    s->d.objInfo = s->flat.objInfo.data();
    s->d.objBox = s->flat.objBox.data();
    s->d.objTris = s->flat.objTris.data();
    s->d.tlasRoot = s->flat.tlasRoot;
    s->d.objectCount = s->flat.objectCount;
    s->d.tlasExtent = s->flat.tlasExtent;
    s->d.tlasNodeCount = s->flat.tlasNodeCount;
    s->d.triangleTest = g_hostsimTriangleTest;
}

// what flatten_scene() made of the scene: [0] worst-case traversal stack entries (TLAS + deepest
// mesh tree), [1] TLAS nodes, [2] deepest tree level, [3] objects
extern "C" void hostsim_flat_info(ora_Scene *s, uint32_t *out4)
{
    out4[0] = s->flat.stackNeed;
    out4[1] = s->flat.tlasNodeCount;
    out4[2] = s->flat.maxDepth;
    out4[3] = s->flat.objectCount;
}

extern "C" int ora_register_material(ora_Scene *s, uint32_t id, const float *albedo,
                                     uint32_t albedoTexture, const float *emission,
                                     uint32_t emissionTexture, float roughness)
{
    if (s->ms.count >= SP_MAX_MATERIALS) return 0;
    sp_Material m = {};
    m.albedo = vec3{albedo[0], albedo[1], albedo[2]};
    m.albedoTexture = albedoTexture;
    m.emission = vec3{emission[0], emission[1], emission[2]};
    m.emissionTexture = emissionTexture;
    m.roughness = roughness;
    uint32_t i = s->ms.count++;
    s->ms.keys[i] = id;
    s->ms.materials[i] = m;
    return 1;
}

extern "C" int ora_register_texture(ora_Scene *s, uint32_t id, const float *pixels,
                                    uint32_t width, uint32_t height)
{
    if (s->ms.imageCount >= SP_MAX_IMAGES) return 0;
    uint32_t i = s->ms.imageCount++;
    s->ms.imageKeys[i] = id;
    s->ms.images[i].pixels = (float *)pixels;
    s->ms.images[i].width = width;
    s->ms.images[i].height = height;
    return 1;
}

extern "C" void ora_set_background(ora_Scene *s, uint32_t id) { s->ms.backgroundMaterialId = id; }

extern "C" void ora_configure_camera(ora_Scene *s, const float *p, const float *q,
                                     float filmDistance, uint32_t width, uint32_t height)
{
    s->plane.width = width;
    s->plane.height = height;
    configure_camera(&s->camera, &s->plane, vec3{p[0], p[1], p[2]}, quat{q[0], q[1], q[2], q[3]},
                     filmDistance);
}

extern "C" uint32_t ora_seed(uint32_t pixelIndex, uint32_t sample, uint32_t frame)
{
    return stream_seed(pixelIndex, sample, frame);
}

static int g_hostsimCull = 1;
extern "C" void hostsim_set_cull(int cull) { g_hostsimCull = cull; }
// 1: traverse with the resumable state machine the production trace kernel runs (single-object entry
// at the start included); 2: the second machine (A/B build -DSPB_TRAV2); 0: the non-resumable walk
static int g_hostsimStepped = 0;
extern "C" void hostsim_set_stepped(int stepped) { g_hostsimStepped = stepped; }

template <bool CULL>
static Hit hs_intersect(const DScene &S, f3 o, f3 d, uint32_t *stack, float *stackT)
{
    return g_hostsimStepped == 1 ? intersect_scene_stepped<CULL>(S, o, d, stack, stackT, nullptr)
           : g_hostsimStepped == 2 ? intersect_scene_stepped2<CULL>(S, o, d, stack, stackT, nullptr)
                                   : intersect_scene<CULL>(S, o, d, stack, stackT, nullptr);
}

template <bool CULL, int STEPPED>
static void render_rows(ora_Scene *s, float *rgba, uint32_t x0, uint32_t y0, uint32_t x1,
                        uint32_t y1, uint32_t spp, uint32_t bounces, uint32_t frame,
                        uint32_t tid, uint32_t threads, uint64_t *m)
{
    uint32_t stack[SPB_STACK_SIZE];
    float stackT[SPB_STACK_SIZE];
    uint32_t width = s->dc.width;
    float weight = 1.0f / (float)spp;
    PathCounters pc = {0, 0, 0};
    uint64_t paths = 0;
    for (uint32_t y = y0 + tid; y < y1; y += threads)
        for (uint32_t x = x0; x < x1; ++x)
        {
            f3 total = mk3(0, 0, 0);
            for (uint32_t sample = 0; sample < spp; ++sample)
            {
                uint32_t rng = stream_seed(x + y * width, sample, frame);
                f3 r = trace_path<0, 0, CULL, STEPPED>(s->d, s->dm, s->dc, x, y, rng, bounces, 10.0f,
                                                       stack, stackT, pc, nullptr);
                total = add3(total, mul3(r, weight));
                paths++;
            }
            float *px = rgba + ((size_t)x + (size_t)y * width) * 4;
            px[0] = total.x; px[1] = total.y; px[2] = total.z; px[3] = 1.0f;
        }
    m[ORA_METRIC_PATHS] = paths;
    m[ORA_METRIC_RAYS] = pc.rays;
    m[ORA_METRIC_HITS] = pc.hits;
    m[ORA_METRIC_MISSES] = pc.misses;
}

extern "C" void ora_render_seeded(ora_Scene *s, float *rgba, uint32_t x0, uint32_t y0,
                                  uint32_t x1, uint32_t y1, uint32_t spp, uint32_t bounces,
                                  uint32_t frame, uint32_t threads, uint64_t *metrics)
{
    refresh(s);
    if (threads == 0) threads = 1;
    std::vector<std::vector<uint64_t>> per(threads, std::vector<uint64_t>(ORA_METRIC_COUNT, 0));
    auto worker = [&](uint32_t tid) {
        uint64_t *out = per[tid].data();
        if (g_hostsimCull && g_hostsimStepped == 1) render_rows<true, 1>(s, rgba, x0, y0, x1, y1, spp, bounces, frame, tid, threads, out);
        else if (g_hostsimCull && g_hostsimStepped == 2) render_rows<true, 2>(s, rgba, x0, y0, x1, y1, spp, bounces, frame, tid, threads, out);
        else if (g_hostsimCull) render_rows<true, 0>(s, rgba, x0, y0, x1, y1, spp, bounces, frame, tid, threads, out);
        else if (g_hostsimStepped == 1) render_rows<false, 1>(s, rgba, x0, y0, x1, y1, spp, bounces, frame, tid, threads, out);
        else if (g_hostsimStepped == 2) render_rows<false, 2>(s, rgba, x0, y0, x1, y1, spp, bounces, frame, tid, threads, out);
        else render_rows<false, 0>(s, rgba, x0, y0, x1, y1, spp, bounces, frame, tid, threads, out);
    };
    std::vector<std::thread> pool;
    for (uint32_t t = 1; t < threads; ++t) pool.emplace_back(worker, t);
    worker(0);
    for (auto &t : pool) t.join();
    if (metrics)
        for (uint32_t t = 0; t < threads; ++t)
            for (int i = 0; i < ORA_METRIC_COUNT; ++i) metrics[i] += per[t][i];
}

extern "C" void ora_path_trace_tile(ora_Scene *s, float *rgba, uint32_t minX, uint32_t minY,
                                    uint32_t maxX, uint32_t maxY, uint32_t spp,
                                    uint32_t bounces, uint32_t *rngState, uint64_t *metrics)
{
    refresh(s);
    uint32_t stack[SPB_STACK_SIZE];
    float stackT[SPB_STACK_SIZE];
    uint32_t width = s->dc.width, height = s->dc.height;
    if (maxX > width) maxX = width;
    if (maxY > height) maxY = height;
    uint32_t rng = *rngState;
    PathCounters pc = {0, 0, 0};
    uint64_t paths = 0;
    float weight = 1.0f / (float)spp;
    for (uint32_t y = minY; y < maxY; ++y)
        for (uint32_t x = minX; x < maxX; ++x)
        {
            f3 total = mk3(0, 0, 0);
            for (uint32_t sample = 0; sample < spp; ++sample)
            {
                f3 r = trace_path<0, 0, true>(s->d, s->dm, s->dc, x, y, rng, bounces, 10.0f, stack,
                                              stackT, pc, nullptr);
                total = add3(total, mul3(r, weight));
                paths++;
            }
            float *px = rgba + ((size_t)x + (size_t)y * width) * 4;
            px[0] = total.x; px[1] = total.y; px[2] = total.z; px[3] = 1.0f;
        }
    *rngState = rng;
    if (metrics)
    {
        metrics[ORA_METRIC_PATHS] += paths;
        metrics[ORA_METRIC_RAYS] += pc.rays;
        metrics[ORA_METRIC_HITS] += pc.hits;
        metrics[ORA_METRIC_MISSES] += pc.misses;
    }
}

extern "C" void ora_primary_hits(ora_Scene *s, int32_t *triId, int32_t *objId, float *tOut,
                                 float *rayDir3, uint32_t sample, uint32_t frame,
                                 uint32_t threads)
{
    refresh(s);
    uint32_t width = s->dc.width, height = s->dc.height;
    if (threads == 0) threads = 1;
    auto worker = [&](uint32_t tid) {
        uint32_t stack[SPB_STACK_SIZE];
        float stackT[SPB_STACK_SIZE];
        for (uint32_t y = tid; y < height; y += threads)
            for (uint32_t x = 0; x < width; ++x)
            {
                uint32_t rng = stream_seed(x + y * width, sample, frame);
                f3 o, d;
                primary_ray(s->dc, x, y, rng, o, d);
                Hit h = g_hostsimCull ? hs_intersect<true>(s->d, o, d, stack, stackT)
                                      : hs_intersect<false>(s->d, o, d, stack, stackT);
                uint32_t index = x + y * width;
                int32_t tri = -1;
                if (h.object >= 0) tri = (int32_t)f2u(s->d.tris[(size_t)h.slot * 3].w);
                if (triId) triId[index] = tri;
                if (objId) objId[index] = h.object;
                if (tOut) tOut[index] = h.t;
                if (rayDir3) { rayDir3[index * 3] = d.x; rayDir3[index * 3 + 1] = d.y; rayDir3[index * 3 + 2] = d.z; }
            }
    };
    std::vector<std::thread> pool;
    for (uint32_t t = 1; t < threads; ++t) pool.emplace_back(worker, t);
    worker(0);
    for (auto &t : pool) t.join();
}

extern "C" void ora_intersect_rays(ora_Scene *s, uint32_t n, const float *origins3,
                                   const float *dirs3, float *out7, int32_t *triId,
                                   int32_t *objId, uint64_t *metrics)
{
    uint32_t stack[SPB_STACK_SIZE];
    float stackT[SPB_STACK_SIZE];
    for (uint32_t i = 0; i < n; ++i)
    {
        f3 o = mk3(origins3[i * 3], origins3[i * 3 + 1], origins3[i * 3 + 2]);
        f3 d = mk3(dirs3[i * 3], dirs3[i * 3 + 1], dirs3[i * 3 + 2]);
        Hit h = g_hostsimCull ? hs_intersect<true>(s->d, o, d, stack, stackT)
                              : hs_intersect<false>(s->d, o, d, stack, stackT);
        float *out = out7 ? out7 + (size_t)i * 7 : nullptr;
        int32_t tri = -1;
        if (h.object >= 0)
        {
            Surface sf = resolve_hit(s->d, h);
            tri = (int32_t)sf.triangle;
            if (out)
            {
                out[0] = h.t;
                memcpy(&out[1], &sf.material, 4);
                out[2] = sf.normal.x; out[3] = sf.normal.y; out[4] = sf.normal.z;
                out[5] = sf.uvx; out[6] = sf.uvy;
            }
        }
        else if (out)
        {
            out[0] = -1.0f;
            out[1] = out[2] = out[3] = out[4] = out[5] = out[6] = 0.0f;
        }
        if (triId) triId[i] = tri;
        if (objId) objId[i] = h.object;
    }
    (void)metrics;
}

// ---------------------------------------------------------------------------------------------
// The conservative box test of the resumable machine (spb_core.cuh slab_wide / trav2_constants)
// against the reference's predicate (slab_fast) on seeded adversarial pairs: flat and point boxes,
// origins on box faces and far outside, directions with components down to 1e-28, grazing rays aimed
// at box corners and edges.  out[0] = pairs, out[1] = pairs the exact test passes, out[2] = pairs the
// wide test passes, out[3] = pairs the exact test passes and the wide one does NOT (must be 0).
extern "C" void hostsim_check_wide_slab(uint32_t seed, uint32_t count, uint64_t *out4)
{
    uint32_t rng = seed | 1u;
    auto uni = [&]() { return rand_unilateral(rng); };
    auto bi = [&]() { return rand_bilateral(rng); };
    out4[0] = out4[1] = out4[2] = out4[3] = 0;
    for (uint32_t i = 0; i < count; ++i)
    {
        const float extent = powf(10.0f, floorf(bi() * 4.0f));           // 1e-4 .. 1e3
        float mn[3], mx[3];
        for (int a = 0; a < 3; ++a)
        {
            float c = bi() * extent * 0.9f, h = uni() * extent * 0.1f;
            uint32_t kind = xorshift32(rng) & 7u;
            if (kind == 0) h = 0.0f;                                       // flat on this axis
            if (kind == 1) h = extent * 1.0e-7f;
            mn[a] = c - h;
            mx[a] = c + h;
            if (mn[a] > mx[a]) { float t = mn[a]; mn[a] = mx[a]; mx[a] = t; }
        }
        f3 o, d;
        uint32_t mode = xorshift32(rng) % 6u;
        float far = (mode == 1) ? 1000.0f : ((mode == 2) ? 30.0f : 2.0f);
        o = mk3(bi() * extent * far, bi() * extent * far, bi() * extent * far);
        if (mode == 3) { o.x = (xorshift32(rng) & 1u) ? mn[0] : mx[0]; }   // on a face
        // aim at a point of the box: a corner, an edge point or an interior point, then perturb by ulps
        float tx = mn[0] + (mx[0] - mn[0]) * (float)(xorshift32(rng) % 3u) * 0.5f;
        float ty = mn[1] + (mx[1] - mn[1]) * (float)(xorshift32(rng) % 3u) * 0.5f;
        float tz = mn[2] + (mx[2] - mn[2]) * (float)(xorshift32(rng) % 3u) * 0.5f;
        d = mk3(tx - o.x, ty - o.y, tz - o.z);
        if (mode == 4) d = mk3(bi(), bi(), bi());
        if (mode == 5) { d.x = bi() * 1.0e-20f; if ((xorshift32(rng) & 3u) == 0) d.y = bi() * 1.0e-28f; }
        d = normalize3(d);
        uint32_t nudge = xorshift32(rng) & 3u;
        if (nudge == 1) d.x = u2f(f2u(d.x) + 1u);
        if (nudge == 2) d.y = u2f(f2u(d.y) - 1u);
        if (!trav2_ray_ok(o, d, extent)) continue;
        f3 inv = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
        float tn;
        bool exact = slab_fast(mn[0], mn[1], mn[2], mx[0], mx[1], mx[2], o, inv, tn);
        Trav2 st;
        trav2_constants(o, inv, extent, st);
        bool wide = slab_wide(mn[0], mn[1], mn[2], mx[0], mx[1], mx[2], st, u2f(0x7F800000u)) < u2f(0x7F800000u);
        out4[0]++;
        out4[1] += exact;
        out4[2] += wide;
        out4[3] += (exact && !wide);
    }
}

// Shortcuts of the wavefront renderer (spb_core.cuh collect_candidates / resolve_from_candidates /
// sky_one_lookup) against the plain evaluation, pixel by pixel, on the host.
//   out[0] pixels, [1] camera rays, [2] pixels whose candidate list fell back, [3] rays whose
//   shortcut result differs from the walk's (t bits, triangle slot or object), [4] of those, rays
//   where only the slot differs at bit-equal t (ties), [5] pixels all of whose samples escape,
//   [6] of those, pixels the one-lookup path settles, [7] of those, pixels whose value differs
//   from the sample loop's, [8] sum of candidate counts.
extern "C" void hostsim_check_shortcuts(ora_Scene *s, uint32_t spp, uint32_t frame, uint32_t threads, uint64_t *out9)
{
    refresh(s);
    if (threads == 0) threads = 1;
    const DCamera &c = s->dc;
    // spread: the host formula of render_wavefront (spb_api.cu)
    double fx = (double)c.filmCenter.x - c.position.x, fy = (double)c.filmCenter.y - c.position.y,
           fz = (double)c.filmCenter.z - c.position.z;
    double dist = sqrt(fx * fx + fy * fy + fz * fz);
    double jitter = fmax(fabs((double)c.halfPixelWidth), fabs((double)c.halfPixelHeight)) +
                    1.2e-7 * (double)(c.width > c.height ? c.width : c.height);
    double pixelAngle = dist > 0.0 ? 2.0 * fmax((double)c.halfFilmWidth / c.width, (double)c.halfFilmHeight / c.height) / dist : 1.0;
    const float spread = (float)(4.0 * jitter * pixelAngle + 1.0e-6);
    const bool single = s->d.objectCount == 1 && s->d.tlasRoot != SPB_REF_EMPTY;
    std::vector<std::vector<uint64_t>> per(threads, std::vector<uint64_t>(9, 0));
    auto worker = [&](uint32_t tid) {
        uint64_t *o9 = per[tid].data();
        uint32_t stack[SPB_STACK_SIZE];
        float stackT[SPB_STACK_SIZE];
        uint32_t list[SPB_CAND_STRIDE];
        const float weight = 1.0f / (float)spp;
        for (uint32_t y = tid; y < c.height; y += threads)
            for (uint32_t x = 0; x < c.width; ++x)
            {
                o9[0]++;
                list[0] = SPB_CAND_FALLBACK;
                if (single) collect_candidates(s->d, c, x, y, list);
                if (list[0] == SPB_CAND_FALLBACK) o9[2]++;
                else o9[8] += list[0];
                bool allMiss = true;
                f3 total = mk3(0, 0, 0);
                for (uint32_t sample = 0; sample < spp; ++sample)
                {
                    uint32_t rng = stream_seed(x + y * c.width, sample, frame);
                    f3 o, d;
                    primary_ray(c, x, y, rng, o, d);
                    Hit walk = intersect_scene_stepped<true>(s->d, o, d, stack, stackT, nullptr);
                    o9[1]++;
                    if (walk.t > 0.0f) allMiss = false;
                    total = add3(total, mul3(miss_radiance<0, 0>(s->dm, neg3(d), 10.0f, nullptr), weight));
                    if (list[0] == SPB_CAND_FALLBACK) continue;
                    // what k_trace<PRIMARY> does with the list
                    Trav st;
                    TravCold cold;
                    trav_begin(s->d, o, d, st, cold);
                    resolve_from_candidates(s->d, list, o, d, st, cold, nullptr);
                    if (st.cur != SPB_NODE_DONE || cold.slow) continue; // the walk takes over
                    Hit fast = trav_result(cold);
                    bool same = f2u(fast.t) == f2u(walk.t) && fast.object == walk.object && (fast.object < 0 || fast.slot == walk.slot);
                    if (!same)
                    {
                        o9[3]++;
                        if (f2u(fast.t) == f2u(walk.t) && fast.object == walk.object) o9[4]++;
                    }
                }
                if (allMiss)
                {
                    o9[5]++;
                    f3 one;
                    if (sky_one_lookup<0, 0>(s->dm, c, x, y, spread, spp, one))
                    {
                        o9[6]++;
                        if (f2u(one.x) != f2u(total.x) || f2u(one.y) != f2u(total.y) || f2u(one.z) != f2u(total.z)) o9[7]++;
                    }
                }
            }
    };
    std::vector<std::thread> pool;
    for (uint32_t t = 1; t < threads; ++t) pool.emplace_back(worker, t);
    worker(0);
    for (auto &t : pool) t.join();
    for (uint32_t t = 0; t < threads; ++t)
        for (int i = 0; i < 9; ++i) out9[i] += per[t][i];
}

// Coverage pass (spb_core.cuh cover_triangle) on the host: out[0] blocks, [1] blocks left
// unmarked, [2] "everything" flag, [3] camera rays (all samples) of pixels in unmarked blocks,
// [4] of those, rays whose walk HITS something (must be 0).
extern "C" void hostsim_check_coverage(ora_Scene *s, uint32_t spp, uint32_t frame, uint32_t threads, uint64_t *out5)
{
    refresh(s);
    if (threads == 0) threads = 1;
    const DCamera &c = s->dc;
    const uint32_t blocksX = (c.width + 7) / 8, blocksY = (c.height + 3) / 4, blocks = blocksX * blocksY;
    std::vector<uint8_t> coverage((size_t)blocks + 1, 0);
    std::vector<uint32_t> objTris = s->flat.objTris;
    if (objTris.empty()) objTris.push_back(0);
    DScene d = s->d;
    d.objTris = objTris.data();
    for (uint64_t i = 0; i < s->flat.instancedTriangles; ++i)
        cover_triangle(d, c, i, 0, 0, c.width, c.height, blocksX, blocksY, coverage.data());
    out5[0] = blocks;
    out5[2] = coverage[blocks];
    std::vector<std::vector<uint64_t>> per(threads, std::vector<uint64_t>(3, 0));
    auto worker = [&](uint32_t tid) {
        uint32_t stack[SPB_STACK_SIZE];
        float stackT[SPB_STACK_SIZE];
        for (uint32_t b = tid; b < blocks; b += threads)
        {
            if (coverage[b] || coverage[blocks]) continue;
            per[tid][0]++;
            for (uint32_t l = 0; l < 32; ++l)
            {
                uint32_t x = (b % blocksX) * 8 + (l & 7u), y = (b / blocksX) * 4 + (l >> 3);
                if (x >= c.width || y >= c.height) continue;
                for (uint32_t sample = 0; sample < spp; ++sample)
                {
                    uint32_t rng = stream_seed(x + y * c.width, sample, frame);
                    f3 o, dir;
                    primary_ray(c, x, y, rng, o, dir);
                    Hit walk = intersect_scene<true>(d, o, dir, stack, stackT, nullptr);
                    per[tid][1]++;
                    if (walk.t > 0.0f) per[tid][2]++;
                }
            }
        }
    };
    std::vector<std::thread> pool;
    for (uint32_t t = 1; t < threads; ++t) pool.emplace_back(worker, t);
    worker(0);
    for (auto &t : pool) t.join();
    for (uint32_t t = 0; t < threads; ++t)
    {
        out5[1] += per[t][0];
        out5[3] += per[t][1];
        out5[4] += per[t][2];
    }
}

// ---- environment pre-processing (spb_cubemap.cuh on the host) ---------------------------------
// Same decomposition as the kernels in spb_cubemap.cu: per-texel state found by the GF(2) jump,
// per-sample terms evaluated independently, folded front to back.
extern "C" void hostsim_create_cube_map(const float *pixels, uint32_t w, uint32_t h, uint32_t faceW,
                                        uint32_t faceH, float *out)
{
    DImage env;
    env.pixels = (const v4f *)pixels;
    env.width = w; env.height = h; env.pad = 0;
    for (uint32_t layer = 0; layer < 6; ++layer)
        for (uint32_t y = 0; y < faceH; ++y)
            for (uint32_t x = 0; x < faceW; ++x)
            {
                v4f t = cube_map_texel<0>(env, layer, x, y, faceW, faceH);
                float *dst = out + (((size_t)layer * faceH + y) * faceW + x) * 4;
                dst[0] = t.x; dst[1] = t.y; dst[2] = t.z; dst[3] = t.w;
            }
}

extern "C" void hostsim_create_irradiance_cube_map(const float *pixels, uint32_t w, uint32_t h,
                                                   uint32_t faceW, uint32_t faceH, uint32_t spp,
                                                   uint32_t sampling, float sampleDelta, float *out)
{
    DImage env;
    env.pixels = (const v4f *)pixels;
    env.width = w; env.height = h; env.pad = 0;
    std::vector<float> phis(irradiance_loop_values(2.0f * SPB_PI, sampleDelta, nullptr, 0));
    std::vector<float> thetas(irradiance_loop_values(0.5f * SPB_PI, sampleDelta, nullptr, 0));
    irradiance_loop_values(2.0f * SPB_PI, sampleDelta, phis.data(), (uint32_t)phis.size());
    irradiance_loop_values(0.5f * SPB_PI, sampleDelta, thetas.data(), (uint32_t)thetas.size());
    std::vector<uint32_t> jumpTexel(32 * 32), jumpSample(32 * 32);
    if (sampling)
    {
        xorshift_build_jump_table(3 * spp, jumpTexel.data());
        xorshift_build_jump_table(3, jumpSample.data());
    }
    const uint32_t sampleCount = sampling ? spp : (uint32_t)(phis.size() * thetas.size());
    std::vector<f3> terms(sampleCount);
    for (uint32_t texel = 0; texel < 6 * faceW * faceH; ++texel)
    {
        uint32_t layer = texel / (faceW * faceH), r = texel % (faceW * faceH);
        uint32_t y = r / faceW, x = r % faceW;
        f3 forward, up, right, tangent, bitangent;
        cube_face_basis(layer, forward, up, right);
        f3 dir = cube_texel_direction(forward, up, right, x, y, faceW, faceH);
        irradiance_frame(up, dir, tangent, bitangent);
        uint32_t texelState = sampling ? xorshift_jump(jumpTexel.data(), texel, 0x45BA12F3u) : 0u;
        for (uint32_t s = sampleCount; s-- > 0;) // any order: the terms are independent
        {
            if (sampling)
            {
                uint32_t rng = xorshift_jump(jumpSample.data(), s, texelState);
                terms[s] = irradiance_random_term<0>(env, dir, rng, 10.0f, 1.0f / (float)spp);
            }
            else
            {
                uint32_t iphi = s / (uint32_t)thetas.size(), itheta = s % (uint32_t)thetas.size();
                terms[s] = irradiance_uniform_term<0>(env, dir, tangent, bitangent, phis[iphi], thetas[itheta], 10.0f);
            }
        }
        f3 sum = mk3(0, 0, 0);
        for (uint32_t s = 0; s < sampleCount; ++s) sum = add3(sum, terms[s]);
        if (!sampling) sum = irradiance_uniform_finish(sum, sampleCount);
        float *dst = out + (size_t)texel * 4;
        dst[0] = sum.x; dst[1] = sum.y; dst[2] = sum.z; dst[3] = 1.0f;
    }
}
