// spb_wavefront.cu -- the wavefront form of the integrator (the production render path).
//
// One frame strip is rendered in passes of S samples per pixel.  A pass is a short sequence of
// kernels over compact device queues (DESIGN.md "Wavefront"):
//
//   k_trace<PRIMARY>   ray generation fused with the first traversal: persistent warps pull
//                      (pixel, sample) items, generate the camera ray (simd_path_tracer.cpp:
//                      216-230) and walk the BVH with the resumable state machine of
//                      spb_core.cuh.  A lane whose ray is finished is retired (hit record +
//                      hit/miss queue entry) and refilled from the queue once fewer than
//                      REFILL_THRESHOLD lanes are still walking -- warp-level compaction and
//                      regeneration, so the SIMT lanes stay full although path lengths differ.
//   k_shade_miss       escaped rays: background material, equirect environment lookup
//                      (sp_material_system.cpp:59-105), fold the path back to the eye
//                      (simd_path_tracer.cpp:113-171).
//   k_shade_hit        surface hits: normal / uv interpolation (sp_scene.cpp:196-216), BSDF terms
//                      and hemisphere sample (math_lib.h:932-946), next ray appended to the
//                      bounce queue (or, on the last bounce, the path folded back).
//   k_trace            the same traversal kernel over the bounce queue.
//   k_accumulate       total += radiance_s * (1/spp) in sample order (simd_path_tracer.cpp:321).
//
// All queues and per-path records live in HBM; nothing returns to the host between kernels (the
// queue lengths are read by the next kernel from device memory).  Arithmetic per path is the
// same sequence of operations as trace_path() in spb_core.cuh, so results are bit-identical to
// the per-pixel kernel and to the oracle in deterministic-math mode.
//
// Compile: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -lineinfo.
#include <map>
#include <cstdio>
#include <cstdlib>

#include "spb_kernels.cuh"

namespace spb {

#define SPB_FULL 0xFFFFFFFFu

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ unsigned lanemask_lt()
{
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// Chunked reservation from a device counter.  Every warp keeps a private range [next, end) of the
// counter's index space in shared memory (`ws`) and hands indices out of it; only when the range
// runs dry does one lane reserve the next `chunk` indices with an atomic.  All warps of the grid
// used to add to the same address once per retire / refill event (about 2 M same-address atomics
// per 33 M-ray launch, serialised in L2); now it is one atomic per `chunk` indices.
// `mask`: lanes that want an index (warp-uniform value); returns this lane's index if it is in mask.
__device__ __forceinline__ unsigned chunk_take(uint32_t *counter, volatile unsigned *ws, unsigned mask,
                                               unsigned chunk)
{
    const unsigned n = __popc(mask), rank = __popc(mask & lanemask_lt());
    const unsigned next = ws[0], end = ws[1];
    const unsigned avail = end - next;
    unsigned idx;
    __syncwarp();
    if (n <= avail)
    {
        idx = next + rank;
        if (lane_id() == 0) ws[0] = next + n;
    }
    else
    {
        unsigned base = 0;
        if (lane_id() == 0) base = atomicAdd(counter, chunk);
        base = __shfl_sync(SPB_FULL, base, 0);
        idx = rank < avail ? next + rank : base + (rank - avail);
        if (lane_id() == 0)
        {
            ws[0] = base + (n - avail);
            ws[1] = base + chunk;
        }
    }
    __syncwarp();
    return idx;
}

// slots of a warp's private state in shared memory
enum { WS_CUR_NEXT = 0, WS_CUR_END, WS_HIT_NEXT, WS_HIT_END, WS_MISS_NEXT, WS_MISS_END, WS_NHITS, WS_NMISSES, WS_COUNT };

__device__ __forceinline__ v4f mk4f(float x, float y, float z, float w)
{
    v4f r;
    r.x = x; r.y = y; r.z = z; r.w = w;
    return r;
}

__device__ __forceinline__ void store_ray(v4f *rays, unsigned slot, f3 o, f3 d, uint32_t rng, uint32_t path)
{
    v4f a, b;
    a.x = o.x; a.y = o.y; a.z = o.z; a.w = u2f(rng);
    b.x = d.x; b.y = d.y; b.z = d.z; b.w = u2f(path);
    rays[(size_t)slot * 2 + 0] = a;
    rays[(size_t)slot * 2 + 1] = b;
}

// Primary work item / path id -> pixel and sample-in-pass.  Items enumerate the pass's blocks (a
// slice of the strip's covered-block list), the 32 pixels of a block row-major, samplesThisPass
// consecutive items per pixel.  Pixels past the strip's edge (block padding) are holes.
__device__ __forceinline__ void item_pixel(const WaveArgs &a, unsigned item, unsigned &x, unsigned &y,
                                           unsigned &sLocal)
{
    unsigned pi = item / a.samplesThisPass;
    sLocal = item - pi * a.samplesThisPass;
    unsigned block = __ldg(a.blockList + (pi >> 5)), l = pi & 31u;
    unsigned by = block / a.blocksX, bx = block - by * a.blocksX;
    x = a.x0 + bx * 8 + (l & 7u);
    y = a.y0 + by * 4 + (l >> 3);
}

// ---------------------------------------------------------------------------------------------
// Candidate triangles of a pixel (single-object scenes): spb_core.cuh collect_candidates() /
// resolve_from_candidates().
__global__ void __launch_bounds__(128)
k_candidates(const __grid_constant__ WaveArgs a, uint32_t *candidates)
{
    const unsigned pixels = a.bandBlocks * 32u;
    for (unsigned pi = blockIdx.x * blockDim.x + threadIdx.x; pi < pixels; pi += gridDim.x * blockDim.x)
    {
        uint32_t *out = candidates + (size_t)pi * SPB_CAND_STRIDE;
        unsigned block = __ldg(a.blockList + (pi >> 5)), l = pi & 31u;
        unsigned by = block / a.blocksX, bx = block - by * a.blocksX;
        unsigned x = a.x0 + bx * 8 + (l & 7u), y = a.y0 + by * 4 + (l >> 3);
        if (x >= a.x1 || y >= a.y1) { out[0] = 0; continue; }
        collect_candidates(a.scene, a.camera, x, y, out);
    }
}

// The rare ray the resumable machine hands back (a reciprocal direction beyond 1e30 or not finite,
// coordinates beyond 1e8): the exact, non-resumable walk.
template <bool CULL>
__device__ __noinline__ Hit slow_intersect(const DScene &S, f3 o, f3 d)
{
    uint32_t stack[SPB_STACK_SIZE];
    float stackT[SPB_STACK_SIZE];
    Hit h = intersect_scene<CULL, false>(S, o, d, stack, stackT, nullptr);
    return h;
}

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void store_terms(v4f *pathTerms, size_t index, const VertexTerms &vt)
{
    v4f a, b;
    a.x = vt.E.x; a.y = vt.E.y; a.z = vt.E.z; a.w = vt.cosine;
    b.x = vt.W.x; b.y = vt.W.y; b.z = vt.W.z; b.w = 0.0f;
    pathTerms[index * 2 + 0] = a;
    pathTerms[index * 2 + 1] = b;
}

__device__ __forceinline__ VertexTerms load_terms(const v4f *pathTerms, size_t index)
{
    v4f a = pathTerms[index * 2 + 0], b = pathTerms[index * 2 + 1];
    VertexTerms vt;
    vt.E = mk3(a.x, a.y, a.z);
    vt.cosine = a.w;
    vt.W = mk3(b.x, b.y, b.z);
    return vt;
}

// radiance of the finished path: the last vertex's radiance, then the stored vertices folded back
// to the eye
__device__ __forceinline__ void finish_path_from(const WaveArgs &a, f3 radiance, uint32_t bounce, uint32_t path)
{
    for (int i = (int)bounce - 1; i >= 0; --i)
        radiance = fold_radiance(load_terms(a.pathTerms, (size_t)i * a.pathCapacity + path), radiance, a.clampValue);
    v4f r;
    r.x = radiance.x; r.y = radiance.y; r.z = radiance.z; r.w = 0.0f;
    a.rad[path] = r;
}

__device__ __forceinline__ void finish_path(const WaveArgs &a, const VertexTerms &last, uint32_t bounce,
                                            uint32_t path)
{
    f3 radiance = fold_radiance(last, mk3(0.0f, 0.0f, 0.0f), a.clampValue);
    for (int i = (int)bounce - 1; i >= 0; --i)
        radiance = fold_radiance(load_terms(a.pathTerms, (size_t)i * a.pathCapacity + path), radiance, a.clampValue);
    v4f r;
    r.x = radiance.x; r.y = radiance.y; r.z = radiance.z; r.w = 0.0f;
    a.rad[path] = r;
}

// cost per tile row (strip rebalancing feedback): SPB_COST_MISS units per escaped ray,
// SPB_COST_HIT per surface hit (a hit costs a shading step and, unless it is the last bounce, a
// far more expensive incoherent traversal); one atomic per distinct row per warp
__device__ __forceinline__ void count_row(const WaveArgs &a, uint32_t path, bool active, unsigned weight)
{
    if (!a.tileRowCost) return;
    unsigned row = 0xFFFFFFFFu;
    if (active) // (an inactive lane's path id may be a hole: not a valid index into the block list)
    {
        unsigned x, y, sLocal;
        item_pixel(a, path, x, y, sLocal);
        row = y / a.tileHeight - a.costRow0;
    }
    unsigned peers = __match_any_sync(SPB_FULL, row);
    if (active && lane_id() == (unsigned)(__ffs(peers) - 1))
        atomicAdd(&a.tileRowCost[row], (unsigned long long)__popc(peers) * weight);
}

// An escaped ray shaded where the trace kernel retires it (WaveArgs::fuseMiss): what k_shade_miss does for one
// entry of the miss queue -- background radiance, the path folded back to the eye, its radiance slot written --
// without the queue entry and the second pass over the ray record.  Out of line: the direction -> texel
// arithmetic (two double-rounded atan2) must not take part in the register allocation of the walk.
__device__ __noinline__ void shade_miss_fused(const WaveArgs &a, float vx, float vy, float vz, uint32_t bounce, uint32_t path)
{
    const DMaterials &M = *a.materials;
    Counters cnt = {0, 0, 0, 0};
    const f3 V = mk3(vx, vy, vz);
    f3 radiance;
    switch ((a.fuseMiss >> 1) & 3u)
    {
    case 0: radiance = miss_radiance<0, 0>(M, V, a.clampValue, &cnt); break;
    case 1: radiance = miss_radiance<1, 0>(M, V, a.clampValue, &cnt); break;
    case 2: radiance = miss_radiance<0, 1>(M, V, a.clampValue, &cnt); break;
    default: radiance = miss_radiance<1, 1>(M, V, a.clampValue, &cnt); break;
    }
    finish_path_from(a, radiance, bounce, path);
    if (a.countStats && cnt.envClamped) atomicAdd(&a.stats[CTR_ENV_CLAMPED], (unsigned long long)cnt.envClamped);
}

#if defined(SPB_TRAV2)
// ---------------------------------------------------------------------------------------------
// A/B BUILD (-DSPB_TRAV2), not the production kernel: the second traversal machine, built on VERDICT
// r01's suggestion that the exact slab predicate is only needed on a leaf's own box.  Measured
// (profiles/r2/README.md): 4.29 G instead of 4.63 G warp instructions per bounce-1 launch of C3, but
// no faster (6.13 vs 6.23 ms: five of its twelve test constants spill, every leaf pays an exact test
// of its own) and 10-13 % slower on the 182-object scene, where the extra leaf step runs at 13 of
// 32 lanes.  What did pay -- entering the object when the ray starts -- was moved into the
// production kernel below.
// Traversal kernel.  Persistent warps; every lane owns one ray at a time and advances it with the
// resumable machine of spb_core.cuh (Trav2: conservative four-FFMA box tests on the way down, the
// reference's exact test on every leaf that comes off the stack).  Each iteration of the inner
// loop the warp runs ONE kind of step -- node visits or leaf work (triangle tests / object entry)
// -- whichever more of its lanes are waiting for, so both code paths execute with most lanes
// active.  When fewer than refillThreshold lanes still have work, finished lanes are retired (hit
// record, hit/miss queue) and refilled from the ray queue.  In single-object scenes the object is
// entered when the ray starts and left when it retires: both with every lane of the warp busy.
// Registers hold what a node step touches (12 test constants, cull distance, current entry, stack
// pointer, object); the rest of a lane's state is a 25-word record in shared memory, word-major so
// that the lanes of a warp hit 32 different banks.
template <bool CULL, bool STATS, int MODE, bool SINGLE>
__global__ void __launch_bounds__(SPB_TRACE_THREADS, SPB_TRACE_MIN_BLOCKS)
k_trace(const __grid_constant__ WaveArgs a, uint32_t bounce)
{
    constexpr bool PRIMARY = MODE == SPB_TRACE_PRIMARY; // (this A/B machine has no eviction / resume modes)
    const unsigned lane = lane_id();
    uint32_t *ctr = a.ctr + bounce * WCTR_STRIDE;
    const unsigned total = PRIMARY ? a.workItems : ctr[WCTR_RAYS];
    uint32_t *cursor = &ctr[WCTR_CURSOR];
    v4f *rays = a.rays[bounce & 1u];
    const bool single = a.scene.objectCount == 1;

    TravEntry stack[SPB_STACK_SIZE];
    __shared__ float recordAll[T2_WORDS][SPB_TRACE_THREADS];
    __shared__ unsigned slotAll[SPB_TRACE_THREADS];
    T2View<SPB_TRACE_THREADS> rec;
    rec.base = &recordAll[0][threadIdx.x];
    unsigned &slot = slotAll[threadIdx.x];
    __shared__ unsigned warpStateAll[SPB_TRACE_THREADS / 32][WS_COUNT];
    volatile unsigned *ws = warpStateAll[threadIdx.x >> 5];
    if (lane < WS_COUNT) ws[lane] = 0;
    __syncwarp();
#if SPB_TLAS_SMEM_NODES > 0
    // north star: "hot top-level nodes staged in shared memory" -- the first SPB_TLAS_SMEM_NODES nodes
    // of the TLAS (breadth-first: the top levels), 9 float4 each.  An A/B build (off by default: the
    // counters that decided it are in profiles/r2/README.md).
    __shared__ v4f tlasNodes[SPB_TLAS_SMEM_NODES * 9];
    const unsigned tlasCount = single ? 0u : (a.scene.tlasNodeCount < SPB_TLAS_SMEM_NODES ? a.scene.tlasNodeCount : SPB_TLAS_SMEM_NODES);
    for (unsigned i = threadIdx.x; i < tlasCount * 8u; i += SPB_TRACE_THREADS)
        tlasNodes[(i >> 3) * 9 + (i & 7u)] = ld4(a.scene.nodes + (size_t)a.scene.tlasRoot * 8 + i);
    __syncthreads();
    const v4f *tlas = tlasCount ? tlasNodes : nullptr;
#else
    const v4f *tlas = nullptr;
    const unsigned tlasCount = 0;
#endif
    // chunk size: a quarter of an even share per warp, so small queues still spread over the GPU
    unsigned chunk = (total / (gridDim.x * (SPB_TRACE_THREADS / 32)) / 4 + 31u) & ~31u;
    chunk = chunk < 32u ? 32u : (chunk > SPB_CHUNK_MAX ? SPB_CHUNK_MAX : chunk);
    Counters cnt = {0, 0, 0, 0};
    Trav2 st;
    st.cur = SPB_NODE_DONE;
    st.sp = 0;
    st.tcull = 0.0f;
    st.object = SPB_T2_NO_OBJECT;
    st.p = st.q = st.cn = st.cf = mk3(0.0f, 0.0f, 0.0f);
    rec.u(T2_SLOW) = 0;
    slot = 0;
    bool have = false;
    bool exhausted = total == 0; // warp-uniform: this warp can get no more work from the queue

    for (;;)
    {
        // ---- retire finished lanes: hit record + queue entry
        const bool finished = have && st.cur == SPB_NODE_DONE;
        if (__any_sync(SPB_FULL, finished))
        {
            Hit h;
            h.t = -1.0f;
            h.slot = 0;
            h.object = -1;
            if (finished)
            {
                if (rec.u(T2_SLOW))
                {
                    f3 wo, wd;
                    trav_world_ray(rays + (size_t)slot * 2, wo, wd);
                    h = slow_intersect<CULL>(a.scene, wo, wd);
                }
                else h = trav2_finish(a.scene, st, rec);
            }
            const bool isHit = finished && h.t > 0.0f;
            const bool isMiss = finished && !isHit;
            const unsigned hitMask = __ballot_sync(SPB_FULL, isHit), missMask = __ballot_sync(SPB_FULL, isMiss);
            // primary hits of a sorted pass are found through hitRec[item], not through a queue
            const bool queueHits = !(PRIMARY && a.sortPrimaryHits);
            unsigned hs = 0, ms = 0;
            if (hitMask && queueHits) hs = chunk_take(&ctr[WCTR_HITS], ws + WS_HIT_NEXT, hitMask, chunk);
            if (missMask) ms = chunk_take(&ctr[WCTR_MISSES], ws + WS_MISS_NEXT, missMask, chunk);
            if (lane == 0)
            {
                ws[WS_NHITS] += __popc(hitMask);
                ws[WS_NMISSES] += __popc(missMask);
            }
            if (isHit)
            {
                unsigned mySlot = slot;
                v4f r;
                r.x = h.t; r.y = u2f(h.slot); r.z = u2f((uint32_t)h.object); r.w = 0.0f;
                a.hitRec[mySlot] = r;
                if (queueHits) a.hitQ[hs] = mySlot;
            }
            if (isMiss)
            {
                a.missQ[ms] = slot;
                if (!queueHits) a.hitRec[slot] = mk4f(-1.0f, 0.0f, 0.0f, 0.0f);
            }
            if (finished) have = false;
        }

        // ---- refill idle lanes from the queue (regeneration)
        if (!exhausted)
        {
            unsigned need = __ballot_sync(SPB_FULL, !have);
            if (need)
            {
                unsigned idx = chunk_take(cursor, ws + WS_CUR_NEXT, need, chunk);
                if (ws[WS_CUR_NEXT] >= total) exhausted = true; // chunks are handed out in order
                if (!have && idx < total)
                {
                    f3 o, d;
                    bool valid = true;
                    if (PRIMARY)
                    {
                        // item -> (pixel in 8x4-block order, sample): neighbouring lanes trace
                        // the samples of one pixel, then the neighbouring pixel.  The reference's
                        // jitter is +-0.5/width of a PIXEL (simd_path_tracer.cpp:222-226), so the
                        // samples of a pixel walk the same nodes: node fetches of a warp coalesce
                        // into one L1 wavefront and its lanes stay together.
                        unsigned x, y, sLocal;
                        item_pixel(a, idx, x, y, sLocal);
                        valid = x < a.x1 && y < a.y1;
                        if (valid)
                        {
                            uint32_t pixelIndex = x + y * a.camera.width;
                            uint32_t rng = stream_seed(pixelIndex, a.firstSample + sLocal, a.frame);
                            primary_ray(a.camera, x, y, rng, o, d);
                            store_ray(rays, idx, o, d, rng, idx); // path id == primary item
                        }
                        else if (a.sortPrimaryHits) a.hitRec[idx] = mk4f(-1.0f, 0.0f, 0.0f, 0.0f);
                    }
                    else
                    {
                        // a slot that stands for a hole of the hit queue it was made from
                        valid = f2u(rays[(size_t)idx * 2 + 1].w) != SPB_QUEUE_HOLE;
                        if (valid) trav_world_ray(rays + (size_t)idx * 2, o, d);
                    }
                    if (valid)
                    {
                        have = true;
                        slot = idx;
                        Counters *c = STATS ? &cnt : nullptr;
                        bool resolved = false;
                        if (PRIMARY && a.candidates)
                        {
                            // camera ray resolved from its pixel's candidate list (spb_core.cuh
                            // resolve_candidates); pixels that fall back walk the tree below
                            if (!trav2_start(a.scene, o, d, st, rec)) resolved = true;
                            else
                                resolved = resolve_from_candidates2(
                                    a.scene, a.candidates + (size_t)(idx / a.samplesThisPass) * SPB_CAND_STRIDE, o, d, st, rec, c);
                        }
                        if (!resolved)
                        {
                            if (single) trav2_begin_single<CULL>(a.scene, o, d, st, rec, stack, c);
                            else trav2_begin(a.scene, o, d, st, rec, stack);
                        }
                    }
                }
            }
        }

        // ---- walk
        // lanes whose walk inside an object has ended leave it together (the exit arithmetic
        // would otherwise run for one or two lanes at a time); single-object scenes never get here
        if (have && st.cur == SPB_NODE_EXIT) trav2_exit<CULL>(a.scene, st, rec, stack);
        unsigned walking = __ballot_sync(SPB_FULL, have && trav_is_walking_ref(st.cur));
        if (!walking)
        {
            if (!__any_sync(SPB_FULL, have) && exhausted) break;
            continue; // everything in flight finished at once
        }
        do
        {
            const bool live = have && trav_is_walking_ref(st.cur);
            const bool wantNode = live && (st.cur & SPB_REF_LEAF) == 0;
            const bool wantLeaf = live && !wantNode;
            unsigned nodeMask = __ballot_sync(SPB_FULL, wantNode);
            unsigned leafMask = walking & ~nodeMask;
            if (__popc(nodeMask) >= __popc(leafMask))
            {
                if (wantNode) trav2_node<CULL>(a.scene, st, rec, stack, STATS ? &cnt : nullptr, tlas, tlasCount);
            }
            else
            {
                if (wantLeaf) trav2_leaf<CULL>(a.scene, st, rec, stack, STATS ? &cnt : nullptr);
            }
            walking = __ballot_sync(SPB_FULL, have && trav_is_walking_ref(st.cur));
        } while (walking && ((unsigned)__popc(walking) >= a.refillThreshold || exhausted));
    }

    // holes: the unused tail of this warp's last chunk of either queue; exact counts
    __syncwarp();
    for (unsigned i = ws[WS_HIT_NEXT] + lane; i < ws[WS_HIT_END]; i += 32) a.hitQ[i] = SPB_QUEUE_HOLE;
    for (unsigned i = ws[WS_MISS_NEXT] + lane; i < ws[WS_MISS_END]; i += 32) a.missQ[i] = SPB_QUEUE_HOLE;
    if (lane == 0)
    {
        if (ws[WS_NHITS]) atomicAdd(&ctr[WCTR_NHITS], ws[WS_NHITS]);
        if (ws[WS_NMISSES]) atomicAdd(&ctr[WCTR_NMISSES], ws[WS_NMISSES]);
    }

    if (STATS)
    {
        unsigned n = __reduce_add_sync(SPB_FULL, cnt.nodeVisits), t = __reduce_add_sync(SPB_FULL, cnt.triangleTests),
                 ob = __reduce_add_sync(SPB_FULL, cnt.objectTests);
        if (lane == 0)
        {
            atomicAdd(&a.stats[CTR_NODE_VISITS], (unsigned long long)n);
            atomicAdd(&a.stats[CTR_TRIANGLE_TESTS], (unsigned long long)t);
            atomicAdd(&a.stats[CTR_OBJECT_TESTS], (unsigned long long)ob);
        }
    }
}
#else // the production kernel: the first machine (exact box tests at every node)
// The trace kernel's stack (spb_core.cuh stack_put / stack_get): entries below SPB_SMEM_STACK in shared memory --
// entry k of thread t at column t of row k, so the 32 lanes of a warp hit 32 different bank pairs whatever their
// stack pointers are -- the rest in the lane's local-memory array.
#ifndef SPB_SMEM_STACK
#define SPB_SMEM_STACK 0
#endif
struct HybridStack
{
    unsigned long long *column; // this thread's column: entry k at column[k * SPB_TRACE_THREADS]
    TravEntry *local;
};
__device__ __forceinline__ void stack_put(const HybridStack &s, int sp, TravEntry e)
{
    if (sp < SPB_SMEM_STACK) s.column[sp * SPB_TRACE_THREADS] = (unsigned long long)e.ref | ((unsigned long long)f2u(e.tnear) << 32);
    else s.local[sp] = e;
}
__device__ __forceinline__ TravEntry stack_get(const HybridStack &s, int sp)
{
    if (sp < SPB_SMEM_STACK)
    {
        unsigned long long raw = s.column[sp * SPB_TRACE_THREADS];
        TravEntry e;
        e.ref = (uint32_t)raw;
        e.tnear = u2f((uint32_t)(raw >> 32));
        return e;
    }
    return load_entry(s.local + sp);
}

// Primary ray of item `idx` resolved from its pixel's candidate list (spb_core.cuh
// resolve_from_candidates); leaves the lane finished, or untouched when the pixel falls back.
#ifndef SPB_VOTE_CLASSES
#define SPB_VOTE_CLASSES 4
#endif
#ifndef SPB_EXIT_WEIGHT
#define SPB_EXIT_WEIGHT 1 // (measured on C5, profiles/r2/s15_*: 182.4 ms per frame with 1, 191.4 with 2, 195.8 with 4; two classes 196.3)
#endif
#ifndef SPB_VOTE_BUSY_EXIT
#define SPB_VOTE_BUSY_EXIT 1 // (C5: 171.8 ms per frame against 181.4 without; profiles/r2/s16_ab_c5_vote_thresholds.txt)
#endif
template <bool CULL>
__device__ __forceinline__ void primary_from_candidates(const WaveArgs &a, unsigned idx, f3 o, f3 d, Trav &st,
                                                        TravCold &cold, Counters *counters)
{
    resolve_from_candidates(a.scene, a.candidates + (size_t)(idx / a.samplesThisPass) * SPB_CAND_STRIDE, o, d, st, cold,
                            counters);
}

// ---------------------------------------------------------------------------------------------
// Traversal kernel.  Persistent warps; every lane owns one ray at a time.  Each iteration of the
// inner loop the warp runs ONE kind of step -- node visits or leaf work (triangle tests / object
// entry) -- whichever more of its lanes are waiting for, so both code paths execute with most
// lanes active.  When fewer than refillThreshold lanes still have work, finished lanes are
// retired (hit record, hit/miss queue) and refilled from the ray queue.
//
// MODE (SPB_TRACE_*): QUEUE -- rays of a bounce queue; PRIMARY -- camera rays generated here; EVICT -- a
// bounce queue walked in packets whose STRAGGLERS are evicted: when fewer than refillThreshold lanes of a
// packet are still walking, their state (a 128-byte continuation record: what a lane needs to go on, its
// few live stack entries included) goes to the continuation buffer and the warp starts a fresh packet
// with all 32 lanes; RESUME -- the second launch of such a step: lanes are refilled from the continuation
// buffer, so the stragglers of many packets walk on together instead of alone in their warps.  The walk
// of a ray is the same sequence of steps either way (spb_core.cuh trav_*), hence the same result.
template <bool CULL, bool STATS, int MODE, bool SINGLE>
__global__ void __launch_bounds__(SPB_TRACE_THREADS, SPB_TRACE_MIN_BLOCKS)
k_trace(const __grid_constant__ WaveArgs a, uint32_t bounce)
{
    constexpr bool PRIMARY = MODE == SPB_TRACE_PRIMARY, EVICT = MODE == SPB_TRACE_EVICT, RESUME = MODE == SPB_TRACE_RESUME;
    const unsigned lane = lane_id();
    uint32_t *ctr = a.ctr + bounce * WCTR_STRIDE;
    // (RESUME: the continuation count of the launch before, clipped to what was written)
    const unsigned total = PRIMARY ? a.workItems : RESUME ? (ctr[WCTR_CONT] < a.contCapacity ? ctr[WCTR_CONT] : a.contCapacity) : ctr[WCTR_RAYS];
    uint32_t *cursor = &ctr[RESUME ? WCTR_CONT_CURSOR : WCTR_CURSOR];
    v4f *rays = a.rays[bounce & 1u];
#if defined(SPB_NO_FUSED_ENTRY)
    constexpr bool single = false; // (A/B build: round 1's behaviour; SINGLE is never set, launch_wave_trace)
#else
    // Single-object scenes (C1-C4): the object is entered when the ray starts and left when it
    // retires, both with every lane of the warp busy, instead of through three more rounds of the
    // step loop (TLAS root, object entry, exit) at whatever lane count the votes give them.
    // Measured on C3: 64.1 ms per frame against 68.2 without (profiles/r2/s6_ab_first_machine_fused_entry.txt).
    // The kernel is compiled once for such scenes (SINGLE; launch_wave_trace picks it): the steps below then
    // carry no object entry, no exit and no stack base (spb_core.cuh trav_pop / trav_leaf).
    constexpr bool single = SINGLE;
#endif

    // per-lane stack in local memory; state that is touched only when an object is entered or
    // left, or the ray retired, lives in shared memory (6 + 1 words per lane, conflict-free stride)
    TravEntry localStack[SPB_STACK_SIZE];
#if SPB_SMEM_STACK > 0
    __shared__ unsigned long long stackRows[SPB_SMEM_STACK][SPB_TRACE_THREADS];
    const HybridStack stack = {&stackRows[0][threadIdx.x], localStack};
#else
    TravEntry *const stack = localStack;
#endif
    __shared__ TravCold coldAll[SPB_TRACE_THREADS];
    __shared__ unsigned slotAll[SPB_TRACE_THREADS];
    TravCold &cold = coldAll[threadIdx.x];
    unsigned &slot = slotAll[threadIdx.x];
    __shared__ unsigned warpStateAll[SPB_TRACE_THREADS / 32][WS_COUNT];
    volatile unsigned *ws = warpStateAll[threadIdx.x >> 5];
    if (lane < WS_COUNT) ws[lane] = 0;
    __syncwarp();
    // chunk size: a quarter of an even share per warp, so small queues still spread over the GPU
    unsigned chunk = (total / (gridDim.x * (SPB_TRACE_THREADS / 32)) / 4 + 31u) & ~31u;
    chunk = chunk < 32u ? 32u : (chunk > SPB_CHUNK_MAX ? SPB_CHUNK_MAX : chunk);
    Counters cnt = {0, 0, 0, 0};
    Trav st;
    st.cur = SPB_NODE_DONE;
    cold.slow = 0;
    slot = 0;
    bool have = false;
    bool exhausted = total == 0; // warp-uniform: this warp can get no more work from the queue

    for (;;)
    {
        // ---- evict the stragglers of a packet (the walk below came back with fewer than refillThreshold
        // lanes still walking): continuation records, one atomic per warp; a lane whose stack is deeper
        // than a record holds stays and walks on beside the next packet
        if (EVICT)
        {
            const bool straggler = have && trav_is_walking(st);
            const unsigned stragglers = __ballot_sync(SPB_FULL, straggler);
            if (stragglers && (unsigned)__popc(stragglers) < a.refillThreshold)
            {
                const bool fits = straggler && st.sp <= (int)SPB_CONT_STACK;
                const unsigned m = __ballot_sync(SPB_FULL, fits);
                if (m)
                {
                    const unsigned leader = (unsigned)__ffs(m) - 1u;
                    unsigned base = 0;
                    if (lane == leader) base = atomicAdd(&ctr[WCTR_CONT], (unsigned)__popc(m));
                    base = __shfl_sync(SPB_FULL, base, leader);
                    const unsigned at = base + __popc(m & lanemask_lt());
                    if (fits && at < a.contCapacity)
                    {
                        v4u *rec = a.cont + (size_t)at * SPB_CONT_QUADS;
                        v4u q;
                        q.x = slot; q.y = st.cur; q.z = f2u(st.tcull); q.w = f2u(st.lT);
                        rec[0] = q;
                        q.x = st.lSlot; q.y = (uint32_t)st.sp | ((uint32_t)((SINGLE ? 0 : st.blasBase) + 1) << 16); q.z = f2u(cold.worldCull); q.w = cold.object;
                        rec[1] = q;
                        q.x = f2u(cold.bT); q.y = cold.bSlot; q.z = (uint32_t)cold.bObject; q.w = 0;
                        rec[2] = q;
                        for (int k = 0; k < st.sp; k += 2)
                        {
                            TravEntry e0 = stack_get(stack, k), e1 = stack_get(stack, k + 1 < st.sp ? k + 1 : k);
                            q.x = e0.ref; q.y = f2u(e0.tnear);
                            q.z = e1.ref; q.w = f2u(e1.tnear);
                            rec[3 + (k >> 1)] = q;
                        }
                        have = false;
                        st.cur = SPB_NODE_DONE;
                    }
                }
            }
        }

        // ---- retire finished lanes: hit record + queue entry
        const bool finished = have && (st.cur == SPB_NODE_DONE || (single && st.cur == SPB_NODE_EXIT));
        if (__any_sync(SPB_FULL, finished))
        {
            Hit h;
            h.t = -1.0f;
            h.slot = 0;
            h.object = -1;
            if (finished)
            {
                if (single && st.cur == SPB_NODE_EXIT)
                {
                    trav_leave(a.scene, st, cold, rays + (size_t)slot * 2);
                    st.cur = SPB_NODE_DONE;
                }
                if (cold.slow)
                {
                    f3 wo, wd;
                    trav_world_ray(rays + (size_t)slot * 2, wo, wd);
                    h = slow_intersect<CULL>(a.scene, wo, wd);
                }
                else h = trav_result(cold);
            }
            const bool isHit = finished && h.t > 0.0f;
            const bool isMiss = finished && !isHit;
            const unsigned hitMask = __ballot_sync(SPB_FULL, isHit), missMask = __ballot_sync(SPB_FULL, isMiss);
            // primary hits of a sorted pass are found through hitRec[item], not through a queue
            const bool queueHits = !(PRIMARY && a.sortPrimaryHits);
            unsigned hs = 0, ms = 0;
            if (hitMask && queueHits) hs = chunk_take(&ctr[WCTR_HITS], ws + WS_HIT_NEXT, hitMask, chunk);
            if (missMask && !a.fuseMiss) ms = chunk_take(&ctr[WCTR_MISSES], ws + WS_MISS_NEXT, missMask, chunk);
            if (lane == 0)
            {
                ws[WS_NHITS] += __popc(hitMask);
                ws[WS_NMISSES] += __popc(missMask);
            }
            if (isHit)
            {
                unsigned mySlot = slot;
                v4f r;
                r.x = h.t; r.y = u2f(h.slot); r.z = u2f((uint32_t)h.object); r.w = 0.0f;
                a.hitRec[mySlot] = r;
                if (queueHits) a.hitQ[hs] = mySlot;
            }
            if (isMiss)
            {
                if (!a.fuseMiss) a.missQ[ms] = slot;
                if (!queueHits) a.hitRec[slot] = mk4f(-1.0f, 0.0f, 0.0f, 0.0f);
            }
            if (a.fuseMiss)
            {
                uint32_t path = 0;
                if (isMiss)
                {
                    const v4f rb = rays[(size_t)slot * 2 + 1]; // (direction, path id)
                    path = f2u(rb.w);
                    shade_miss_fused(a, -rb.x, -rb.y, -rb.z, bounce, path);
                }
                if (missMask) count_row(a, path, isMiss, SPB_COST_MISS);
            }
            if (finished) have = false;
        }

        // ---- refill idle lanes from the queue (regeneration)
        if (!exhausted)
        {
            unsigned need = __ballot_sync(SPB_FULL, !have);
            if (need)
            {
                unsigned idx = chunk_take(cursor, ws + WS_CUR_NEXT, need, chunk);
                if (ws[WS_CUR_NEXT] >= total) exhausted = true; // chunks are handed out in order
                if (RESUME)
                {
                    if (!have && idx < total)
                    {
                        // a continuation: the state the EVICT launch parked; the ray in the space it was
                        // being walked in is recomputed with the arithmetic that produced it the first time
                        const v4u *rec = a.cont + (size_t)idx * SPB_CONT_QUADS;
                        v4u q0 = rec[0], q1 = rec[1], q2 = rec[2];
                        slot = q0.x;
                        st.cur = q0.y; st.tcull = u2f(q0.z); st.lT = u2f(q0.w);
                        st.lSlot = q1.x; st.sp = (int)(q1.y & 0xFFFFu); st.blasBase = (int)(q1.y >> 16) - 1;
                        cold.worldCull = u2f(q1.z); cold.object = q1.w;
                        cold.bT = u2f(q2.x); cold.bSlot = q2.y; cold.bObject = (int32_t)q2.z; cold.slow = 0;
                        for (int k = 0; k < st.sp; k += 2)
                        {
                            v4u q = rec[3 + (k >> 1)];
                            TravEntry e0, e1;
                            e0.ref = q.x; e0.tnear = u2f(q.y);
                            e1.ref = q.z; e1.tnear = u2f(q.w);
                            stack_put(stack, k, e0);
                            if (k + 1 < st.sp) stack_put(stack, k + 1, e1);
                        }
                        f3 wo, wd;
                        trav_world_ray(rays + (size_t)slot * 2, wo, wd);
                        if (SINGLE || st.blasBase >= 0)
                        {
                            m4 invModel = load_m4(a.scene.objInv + (size_t)cold.object * 4);
                            st.o = xform(invModel, wo, 1.0f);
                            st.d = normalize3(xform(invModel, wd, 0.0f));
                        }
                        else { st.o = wo; st.d = wd; }
                        st.inv = mk3(1.0f / st.d.x, 1.0f / st.d.y, 1.0f / st.d.z);
                        have = true;
                    }
                }
                else if (!have && idx < total)
                {
                    f3 o, d;
                    bool valid = true;
                    if (PRIMARY)
                    {
                        // item -> (pixel in 8x4-block order, sample): neighbouring lanes trace
                        // the samples of one pixel, then the neighbouring pixel.  The reference's
                        // jitter is +-0.5/width of a PIXEL (simd_path_tracer.cpp:222-226), so the
                        // samples of a pixel walk the same nodes: node fetches of a warp coalesce
                        // into one L1 wavefront and its lanes stay together.
                        unsigned x, y, sLocal;
                        item_pixel(a, idx, x, y, sLocal);
                        valid = x < a.x1 && y < a.y1;
                        if (valid)
                        {
                            uint32_t pixelIndex = x + y * a.camera.width;
                            uint32_t rng = stream_seed(pixelIndex, a.firstSample + sLocal, a.frame);
                            primary_ray(a.camera, x, y, rng, o, d);
                            store_ray(rays, idx, o, d, rng, idx); // path id == primary item
                        }
                        else if (a.sortPrimaryHits) a.hitRec[idx] = mk4f(-1.0f, 0.0f, 0.0f, 0.0f);
                    }
                    else
                    {
                        // a slot that stands for a hole of the hit queue it was made from
                        valid = f2u(rays[(size_t)idx * 2 + 1].w) != SPB_QUEUE_HOLE;
                        if (valid) trav_world_ray(rays + (size_t)idx * 2, o, d);
                    }
                    if (valid)
                    {
                        have = true;
                        slot = idx;
                        if (single && !(PRIMARY && a.candidates)) trav_begin_single<CULL>(a.scene, o, d, st, cold, STATS ? &cnt : nullptr);
                        else
                        {
                            trav_begin(a.scene, o, d, st, cold);
                            if (PRIMARY && a.candidates)
                            {
                                uint32_t before = st.cur;
                                primary_from_candidates<CULL>(a, idx, o, d, st, cold, STATS ? &cnt : nullptr);
                                // a pixel that falls back to the walk
                                if (single && st.cur == before && st.cur != SPB_NODE_DONE && !cold.slow)
                                    trav_begin_single<CULL>(a.scene, o, d, st, cold, STATS ? &cnt : nullptr);
                            }
                        }
                    }
                }
            }
        }

        // ---- walk
        // lanes whose walk inside an object has ended leave it together (the exit arithmetic
        // would otherwise run for one or two lanes at a time)
        if (!single && have && st.cur == SPB_NODE_EXIT)
            trav_exit<CULL>(a.scene, st, cold, rays + (size_t)slot * 2, stack);
        unsigned walking = __ballot_sync(SPB_FULL, have && trav_is_walking(st));
        if (!walking)
        {
            if (!__any_sync(SPB_FULL, have) && exhausted) break;
            continue; // everything in flight finished at once
        }
#if SPB_EARLY_FETCH == 2
        // The node a lane wants to visit is fetched BEFORE the warp votes on the kind of the next step, so the
        // L1 round trip runs under the vote and the branch instead of in front of the first slab test; in an
        // iteration that runs a leaf step the data is dropped (and fetched again next time: an L1 hit).
        do
        {
            const bool live = have && trav_is_walking(st);
            const bool wantNode = live && trav_is_node(st);
            const bool wantLeaf = live && !wantNode;
            NodeData nd;
            trav_node_fetch(a.scene, wantNode, st.cur, nd);
            unsigned nodeMask = __ballot_sync(SPB_FULL, wantNode);
            unsigned leafMask = walking & ~nodeMask;
            if (__popc(nodeMask) >= __popc(leafMask))
            {
                if (wantNode) trav_node_apply<CULL, SINGLE>(a.scene, st, stack, STATS ? &cnt : nullptr, nd);
            }
            else
            {
                if (wantLeaf)
                    trav_leaf<CULL, SINGLE>(a.scene, st, cold, rays + (size_t)slot * 2, stack, STATS ? &cnt : nullptr);
            }
            walking = __ballot_sync(SPB_FULL, have && trav_is_walking(st));
        } while (walking && ((unsigned)__popc(walking) >= a.refillThreshold || (exhausted && !EVICT)));
#elif SPB_EARLY_FETCH == 1
        // The node a lane will visit next is fetched the moment it is known -- at the end of the step that
        // made it the current entry -- so the L1 round trip runs under the two votes and the branch to the
        // next step instead of in front of its first slab test.  A lane that wants a node while the warp runs
        // a leaf step fetches it again afterwards (an L1 hit).
        bool live = have && trav_is_walking(st);
        bool wantNode = live && trav_is_node(st);
        NodeData nd;
        trav_node_fetch(a.scene, wantNode, st.cur, nd);
        unsigned nodeMask = __ballot_sync(SPB_FULL, wantNode);
        do
        {
            if (__popc(nodeMask) >= __popc(walking & ~nodeMask))
            {
                if (wantNode) trav_node_apply<CULL, SINGLE>(a.scene, st, stack, STATS ? &cnt : nullptr, nd);
            }
            else
            {
                if (live && !wantNode)
                    trav_leaf<CULL, SINGLE>(a.scene, st, cold, rays + (size_t)slot * 2, stack, STATS ? &cnt : nullptr);
            }
            live = have && trav_is_walking(st);
            wantNode = live && trav_is_node(st);
            trav_node_fetch(a.scene, wantNode, st.cur, nd);
            walking = __ballot_sync(SPB_FULL, live);
            nodeMask = __ballot_sync(SPB_FULL, wantNode);
        } while (walking && ((unsigned)__popc(walking) >= a.refillThreshold || (exhausted && !EVICT)));
#else
#if SPB_VOTE_CLASSES == 4
        // Several objects: four kinds of step -- node visit, triangle test, object entry, object exit -- and the
        // warp runs the one most of its lanes wait for (an exit counts SPB_EXIT_WEIGHT times: it is the
        // cheapest step and gives a lane back to the other three).  With two classes (node / leaf) an entry and
        // a triangle test in the same leaf step ran one after the other, and a lane whose object had ended idled
        // until the warp went round the outer loop.
        if (!SINGLE)
        {
            do
            {
                const bool live = have && trav_is_walking(st);
                const bool wantNode = live && trav_is_node(st);
                const bool inObject = st.blasBase >= 0;
                const bool wantTri = live && !wantNode && inObject;
                const bool wantEnter = live && !wantNode && !inObject;
                const bool wantExit = have && st.cur == SPB_NODE_EXIT;
                const int n = __popc(__ballot_sync(SPB_FULL, wantNode)), t = __popc(__ballot_sync(SPB_FULL, wantTri)),
                          e = __popc(__ballot_sync(SPB_FULL, wantEnter)), x = __popc(__ballot_sync(SPB_FULL, wantExit)) * SPB_EXIT_WEIGHT;
                // (the majority; voting for the step that leaves the fewest lane-slots idle, cost x (32 - lanes), which a
                // host lane model favoured by 2 %, measured 1.7 % slower on C3 and 22 % slower on C5: profiles/r2/s17_*)
                const bool pickNode = n >= t && n >= e && n >= x, pickTri = !pickNode && t >= e && t >= x,
                           pickEnter = !pickNode && !pickTri && e >= x;
                if (pickNode)
                {
                    if (wantNode) trav_node<CULL, false>(a.scene, st, stack, STATS ? &cnt : nullptr);
                }
                else if (pickTri)
                {
                    if (wantTri) trav_leaf<CULL, false>(a.scene, st, cold, rays + (size_t)slot * 2, stack, STATS ? &cnt : nullptr);
                }
                else if (pickEnter)
                {
                    if (wantEnter) trav_leaf<CULL, false>(a.scene, st, cold, rays + (size_t)slot * 2, stack, STATS ? &cnt : nullptr);
                }
                else
                {
                    if (wantExit) trav_exit<CULL>(a.scene, st, cold, rays + (size_t)slot * 2, stack);
                }
#if SPB_VOTE_BUSY_EXIT
                walking = __ballot_sync(SPB_FULL, have && st.cur <= SPB_NODE_EXIT); // (a lane about to leave its object has work in this loop)
#else
                walking = __ballot_sync(SPB_FULL, have && trav_is_walking(st));
#endif
            } while (walking && ((unsigned)__popc(walking) >= a.refillThreshold || (exhausted && !EVICT)));
            continue;
        }
#endif
        do
        {
            const bool live = have && trav_is_walking(st);
            const bool wantNode = live && trav_is_node(st);
            const bool wantLeaf = live && !wantNode;
            unsigned nodeMask = __ballot_sync(SPB_FULL, wantNode);
            unsigned leafMask = walking & ~nodeMask;
            const bool pickNode = __popc(nodeMask) >= __popc(leafMask);
            if (pickNode)
            {
                if (wantNode) trav_node<CULL, SINGLE>(a.scene, st, stack, STATS ? &cnt : nullptr);
            }
            else
            {
                if (wantLeaf)
                    trav_leaf<CULL, SINGLE>(a.scene, st, cold, rays + (size_t)slot * 2, stack, STATS ? &cnt : nullptr);
            }
            walking = __ballot_sync(SPB_FULL, have && trav_is_walking(st));
        } while (walking && ((unsigned)__popc(walking) >= a.refillThreshold || (exhausted && !EVICT)));
#endif
    }


    // holes: the unused tail of this warp's last chunk of either queue; exact counts
    __syncwarp();
    for (unsigned i = ws[WS_HIT_NEXT] + lane; i < ws[WS_HIT_END]; i += 32) a.hitQ[i] = SPB_QUEUE_HOLE;
    for (unsigned i = ws[WS_MISS_NEXT] + lane; i < ws[WS_MISS_END]; i += 32) a.missQ[i] = SPB_QUEUE_HOLE;
    if (lane == 0)
    {
        if (ws[WS_NHITS]) atomicAdd(&ctr[WCTR_NHITS], ws[WS_NHITS]);
        if (ws[WS_NMISSES]) atomicAdd(&ctr[WCTR_NMISSES], ws[WS_NMISSES]);
    }

    if (STATS)
    {
        unsigned n = __reduce_add_sync(SPB_FULL, cnt.nodeVisits), t = __reduce_add_sync(SPB_FULL, cnt.triangleTests),
                 ob = __reduce_add_sync(SPB_FULL, cnt.objectTests);
        if (lane == 0)
        {
            atomicAdd(&a.stats[CTR_NODE_VISITS], (unsigned long long)n);
            atomicAdd(&a.stats[CTR_TRIANGLE_TESTS], (unsigned long long)t);
            atomicAdd(&a.stats[CTR_OBJECT_TESTS], (unsigned long long)ob);
        }
    }
}

#endif // SPB_TRAV2

// Escaped rays: one thread per entry of the miss queue.  Every record a miss touches -- its queue entry, its ray,
// the vertex terms of its path, one texel of the environment map, its radiance slot -- is a scattered 16- or
// 32-byte access behind the one before (profiles/r2/s8_shade_miss_ncu_full.txt: 33 warps per issue waiting on
// memory, DRAM at 56 %).  Measured and NOT adopted (profiles/r2/README.md): running the chain one and two items
// ahead with prefetch.global.L2 (+0.1 ... +1.1 ms per frame), 32 / 36 / 56-64 registers instead of 40 (+0.9 ...
// +1.8 ms); adopted: a grid of exactly the resident CTAs (launch_wave_shade).
template <int MATH, int ENVFILTER>
__global__ void __launch_bounds__(256)
k_shade_miss(const __grid_constant__ WaveArgs a, uint32_t bounce)
{
    const uint32_t *ctr = a.ctr + bounce * WCTR_STRIDE;
    const unsigned total = ctr[WCTR_MISSES];
    const v4f *rays = a.rays[bounce & 1u];
    const DMaterials &M = *a.materials;
    Counters cnt = {0, 0, 0, 0};
    const unsigned stride = gridDim.x * blockDim.x;
    const unsigned rounds = (total + stride - 1) / stride;
    for (unsigned k = 0; k < rounds; ++k)
    {
        unsigned i = k * stride + blockIdx.x * blockDim.x + threadIdx.x;
        bool active = i < total;
        uint32_t path = 0;
        unsigned slot = active ? a.missQ[i] : SPB_QUEUE_HOLE;
        active = slot != SPB_QUEUE_HOLE;
        if (active)
        {
            v4f rb = rays[(size_t)slot * 2 + 1];
            path = f2u(rb.w);
            f3 V = neg3(mk3(rb.x, rb.y, rb.z));
            finish_path_from(a, miss_radiance<MATH, ENVFILTER>(M, V, a.clampValue, &cnt), bounce, path);
        }
        count_row(a, path, active, SPB_COST_MISS);
    }
    if (a.countStats)
    {
        unsigned e = __reduce_add_sync(SPB_FULL, cnt.envClamped);
        if (lane_id() == 0 && e) atomicAdd(&a.stats[CTR_ENV_CLAMPED], (unsigned long long)e);
    }
}

// One surface hit (simd_path_tracer.cpp:262-300): barycentrics, surface attributes, hemisphere
// sample and BSDF terms; on the last bounce the path is folded back, otherwise its vertex terms
// are stored and the next ray returned.
template <int MATH, int ENVFILTER>
__device__ __forceinline__ void shade_hit_one(const WaveArgs &a, const DMaterials &M, const v4f *rays, uint32_t bounce,
                                              unsigned slot, bool last, Counters &cnt, f3 &no, f3 &nd,
                                              uint32_t &rng, uint32_t &path)
{
    v4f ra = rays[(size_t)slot * 2 + 0], rb = rays[(size_t)slot * 2 + 1];
    v4f hr = a.hitRec[slot];
    f3 o = mk3(ra.x, ra.y, ra.z), d = mk3(rb.x, rb.y, rb.z);
    rng = f2u(ra.w);
    path = f2u(rb.w);
    Hit hit;
    hit.t = hr.x; hit.slot = f2u(hr.y); hit.object = (int32_t)f2u(hr.z);
    hit_barycentrics(a.scene, o, d, hit);
    // simd_path_tracer.cpp:268-289
    Surface sf = resolve_hit(a.scene, hit);
    f3 V = neg3(d);
    f3 P = add3(o, mul3(d, hit.t));
    f3 L = random_hemisphere<MATH>(sf.normal, rng);
    VertexTerms vt = vertex_terms<MATH, ENVFILTER>(M, sf.material, L, sf.normal, V, sf.uvx, sf.uvy, &cnt);
    if (last)
    {
        finish_path(a, vt, bounce, path);
    }
    else
    {
        store_terms(a.pathTerms, (size_t)bounce * a.pathCapacity + path, vt);
        no = add3(P, mul3(sf.normal, 0.0001f));
        nd = L;
    }
}

template <int MATH, int ENVFILTER>
__global__ void __launch_bounds__(256)
k_shade_hit(const __grid_constant__ WaveArgs a, uint32_t bounce)
{
    uint32_t *ctr = a.ctr + bounce * WCTR_STRIDE;
    const unsigned total = ctr[WCTR_HITS];
    const v4f *rays = a.rays[bounce & 1u];
    v4f *nextRays = a.rays[(bounce + 1u) & 1u];
    const DMaterials &M = *a.materials;
    Counters cnt = {0, 0, 0, 0};
    const bool last = bounce + 1 >= a.bounces;
    const unsigned stride = gridDim.x * blockDim.x;
    const unsigned rounds = (total + stride - 1) / stride;
    for (unsigned k = 0; k < rounds; ++k)
    {
        unsigned i = k * stride + blockIdx.x * blockDim.x + threadIdx.x;
        const bool inQueue = i < total;
        unsigned slot = inQueue ? a.hitQ[i] : SPB_QUEUE_HOLE;
        bool active = slot != SPB_QUEUE_HOLE;
        f3 no = mk3(0, 0, 0), nd = mk3(0, 0, 0);
        uint32_t rng = 0, path = SPB_QUEUE_HOLE;
        if (active) shade_hit_one<MATH, ENVFILTER>(a, M, rays, bounce, slot, last, cnt, no, nd, rng, path);
        // the next bounce's ray queue is the hit queue, slot for slot (holes stay holes: path id
        // SPB_QUEUE_HOLE), so no counter is touched here
        if (!last && inQueue) store_ray(nextRays, i, no, nd, rng, path);
        count_row(a, path, active, SPB_COST_HIT);
    }
    if (!last && blockIdx.x == 0 && threadIdx.x == 0) ctr[WCTR_STRIDE + WCTR_RAYS] = total;
    if (a.countStats)
    {
        unsigned e = __reduce_add_sync(SPB_FULL, cnt.envClamped);
        if (lane_id() == 0 && e) atomicAdd(&a.stats[CTR_ENV_CLAMPED], (unsigned long long)e);
    }
}

// Direction bin of a bounce ray: the octahedral map of the direction, 4 bits per coordinate,
// interleaved (Morton order), so that consecutive bins are neighbouring cones.
#ifndef SPB_DIR_BINS
#define SPB_DIR_BINS 256u
#endif
__device__ __forceinline__ unsigned direction_bin(f3 d)
{
#if SPB_DIR_BINS == 64u
    unsigned octant = (d.x < 0.0f ? 1u : 0u) | (d.y < 0.0f ? 2u : 0u) | (d.z < 0.0f ? 4u : 0u);
    const float c = 0.57735f;
    unsigned shape = (fabsf(d.x) > c ? 1u : 0u) | (fabsf(d.y) > c ? 2u : 0u) | (fabsf(d.z) > c ? 4u : 0u);
    return octant * 8u + shape;
#else
    float n = fabsf(d.x) + fabsf(d.y) + fabsf(d.z);
    float inv = n > 0.0f ? __fdividef(1.0f, n) : 0.0f;
    float u = d.x * inv, v = d.y * inv;
    if (d.z < 0.0f)
    {
        float fu = (1.0f - fabsf(v)) * (u >= 0.0f ? 1.0f : -1.0f);
        float fv = (1.0f - fabsf(u)) * (v >= 0.0f ? 1.0f : -1.0f);
        u = fu; v = fv;
    }
    int iu = (int)((u * 0.5f + 0.5f) * 16.0f), iv = (int)((v * 0.5f + 0.5f) * 16.0f);
    unsigned qu = (unsigned)(iu < 0 ? 0 : (iu > 15 ? 15 : iu)), qv = (unsigned)(iv < 0 ? 0 : (iv > 15 ? 15 : iv));
    // interleave the two 4-bit coordinates
    qu = (qu | (qu << 2)) & 0x33u; qu = (qu | (qu << 1)) & 0x55u;
    qv = (qv | (qv << 2)) & 0x33u; qv = (qv | (qv << 1)) & 0x55u;
    return qu | (qv << 1);
#endif
}

// Primary hits, tile by tile.  A tile is SPB_SORT_TILE consecutive primary items -- the samples of
// the pixels of one 8x4 block when samplesThisPass = 64 -- so its bounce rays leave (almost) one
// surface point per pixel and differ in direction only.  The CTA shades the tile's hits, bins the
// new rays by direction and writes them bin by bin into the tile's slots of the next ray queue:
// the trace kernel then hands a warp 32 rays with neighbouring origins AND similar directions,
// which walk the same nodes (coalesced node fetches, lanes that finish together).  Slots the tile
// does not fill are holes.  The order of rays in a queue has no influence on any result.
template <int MATH, int ENVFILTER>
__global__ void __launch_bounds__(256)
k_shade_hit_tiles(const __grid_constant__ WaveArgs a, uint32_t bounce)
{
    __shared__ unsigned binCount[SPB_DIR_BINS], binStart[SPB_DIR_BINS];
    __shared__ unsigned meta[SPB_SORT_TILE];
    // bounce 0: the tile's items are primary items (results in hitRec by item); later bounces:
    // consecutive entries of the hit queue
    uint32_t *ctr = a.ctr + bounce * WCTR_STRIDE;
    const v4f *rays = a.rays[bounce & 1u];
    v4f *nextRays = a.rays[(bounce + 1u) & 1u];
    const DMaterials &M = *a.materials;
    Counters cnt = {0, 0, 0, 0};
    const bool byItem = bounce == 0;
    const unsigned entries = byItem ? a.workItems : ctr[WCTR_HITS];
    const unsigned tiles = (entries + SPB_SORT_TILE - 1) / SPB_SORT_TILE;
    for (unsigned tile = blockIdx.x; tile < tiles; tile += gridDim.x)
    {
        const unsigned base = tile * SPB_SORT_TILE;
        for (unsigned b = threadIdx.x; b < SPB_DIR_BINS; b += 256) binCount[b] = 0;
        __syncthreads();
        for (unsigned k = 0; k < SPB_SORT_TILE / 256; ++k)
        {
            const unsigned local = k * 256 + threadIdx.x, item = base + local;
            unsigned slot = SPB_QUEUE_HOLE;
            if (item < entries) slot = byItem ? (a.hitRec[item].x > 0.0f ? item : SPB_QUEUE_HOLE) : a.hitQ[item];
            bool active = slot != SPB_QUEUE_HOLE;
            f3 no = mk3(0, 0, 0), nd = mk3(0, 0, 0);
            uint32_t rng = 0, path = SPB_QUEUE_HOLE;
            unsigned m = 0xFFFFFFFFu;
            if (active)
            {
                shade_hit_one<MATH, ENVFILTER>(a, M, rays, bounce, slot, false, cnt, no, nd, rng, path);
                unsigned bin = direction_bin(nd);
                m = (bin << 16) | atomicAdd(&binCount[bin], 1u);
                store_ray(a.stage, item, no, nd, rng, path);
            }
            meta[local] = m;
            count_row(a, path, active, SPB_COST_HIT);
        }
        __syncthreads();
        if (threadIdx.x < 32)
        {
            // exclusive prefix sum over the bins: SPB_DIR_BINS / 32 consecutive bins per lane
            const unsigned per = SPB_DIR_BINS / 32u, b0 = threadIdx.x * per;
            unsigned mine = 0;
            for (unsigned b = 0; b < per; ++b) mine += binCount[b0 + b];
            unsigned incl = mine;
            for (unsigned dlt = 1; dlt < 32; dlt <<= 1)
            {
                unsigned v = __shfl_up_sync(SPB_FULL, incl, dlt);
                if (threadIdx.x >= dlt) incl += v;
            }
            unsigned at = incl - mine;
            for (unsigned b = 0; b < per; ++b) { binStart[b0 + b] = at; at += binCount[b0 + b]; }
            __syncwarp();
            if (threadIdx.x == 31) binCount[0] = incl; // rays of the tile
        }
        __syncthreads();
        const unsigned filled = binCount[0];
        for (unsigned k = 0; k < SPB_SORT_TILE / 256; ++k)
        {
            const unsigned local = k * 256 + threadIdx.x;
            const unsigned m = meta[local];
            if (m != 0xFFFFFFFFu)
            {
                const unsigned dst = base + binStart[m >> 16] + (m & 0xFFFFu);
                nextRays[(size_t)dst * 2 + 0] = a.stage[(size_t)(base + local) * 2 + 0];
                nextRays[(size_t)dst * 2 + 1] = a.stage[(size_t)(base + local) * 2 + 1];
            }
            if (local >= filled) nextRays[(size_t)(base + local) * 2 + 1] = mk4f(0.0f, 0.0f, 0.0f, u2f(SPB_QUEUE_HOLE));
        }
        __syncthreads();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) ctr[WCTR_STRIDE + WCTR_RAYS] = tiles * SPB_SORT_TILE;
    if (a.countStats) // (k_shade_hit_tiles)
    {
        unsigned e = __reduce_add_sync(SPB_FULL, cnt.envClamped);
        if (lane_id() == 0 && e) atomicAdd(&a.stats[CTR_ENV_CLAMPED], (unsigned long long)e);
    }
}

// total += radiance_s * (1 / spp), s in order (simd_path_tracer.cpp:216-321, 335-339)
__global__ void __launch_bounds__(256)
k_accumulate(const __grid_constant__ WaveArgs a)
{
    const float weight = 1.0f / (float)a.spp;
    const unsigned pixels = a.bandBlocks * 32u;
    for (unsigned pi = blockIdx.x * blockDim.x + threadIdx.x; pi < pixels; pi += gridDim.x * blockDim.x)
    {
        unsigned block = __ldg(a.blockList + (pi >> 5)), l = pi & 31u;
        unsigned by = block / a.blocksX, bx = block - by * a.blocksX;
        unsigned x = a.x0 + bx * 8 + (l & 7u), y = a.y0 + by * 4 + (l >> 3);
        if (x >= a.x1 || y >= a.y1) continue;
        size_t pixel = (size_t)x + (size_t)y * a.camera.width;
        const v4f *rad = a.rad + (size_t)pi * a.samplesThisPass;
        f3 total = mk3(0.0f, 0.0f, 0.0f);
        if (a.firstSample != 0)
        {
            v4f prev = a.out[pixel];
            total = mk3(prev.x, prev.y, prev.z);
        }
        for (unsigned s = 0; s < a.samplesThisPass; ++s)
        {
            v4f r = rad[s];
            total = add3(total, mul3(mk3(r.x, r.y, r.z), weight));
        }
        v4f o;
        o.x = total.x; o.y = total.y; o.z = total.z; o.w = 1.0f;
        a.out[pixel] = o;
    }
}

// ---------------------------------------------------------------------------------------------
// Coverage pass.  One thread per instanced triangle: object-space vertices -> world (model
// matrix) -> film (sp_CalculateFilmPositions inverted, simd_path_tracer.cpp:38-63) in double; the
// pixel bounding box of the three projections, padded by 2 pixels, marks the 8x4 blocks it touches.
// The padding is four orders of magnitude above the jitter (+-0.5/width of a pixel,
// simd_path_tracer.cpp:222-226) and any rounding of ray generation, so a camera ray through an
// unmarked block misses every triangle with a margin no float intersection test can bridge: its
// closest hit is "none", whatever the box tests on the way would have said.
__global__ void __launch_bounds__(256)
k_cover(const __grid_constant__ WaveArgs a, unsigned long long triangles, uint8_t *coverage)
{
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < triangles;
         i += (unsigned long long)gridDim.x * blockDim.x)
        cover_triangle(a.scene, a.camera, i, a.x0, a.y0, a.x1, a.y1, a.blocksX, a.blocksY, coverage);
}

// Row-major list of the marked blocks (one CTA: a chunked prefix sum); coverage[] becomes the
// final mask (the "everything" flag folded in).
__global__ void __launch_bounds__(1024)
k_list_blocks(unsigned blocks, uint8_t *coverage, uint32_t *blockList, uint32_t *listCount)
{
    __shared__ unsigned sums[1024];
    __shared__ unsigned listHash;
    if (threadIdx.x == 0) listHash = 0;
    const bool all = coverage[blocks] != 0;
    const unsigned per = (blocks + 1023u) / 1024u;
    const unsigned b0 = threadIdx.x * per, b1 = b0 + per < blocks ? b0 + per : blocks;
    unsigned n = 0;
    for (unsigned b = b0; b < b1; ++b)
    {
        if (all) coverage[b] = 1;
        n += coverage[b] != 0;
    }
    sums[threadIdx.x] = n;
    __syncthreads();
    for (unsigned d = 1; d < 1024; d <<= 1)
    {
        unsigned v = threadIdx.x >= d ? sums[threadIdx.x - d] : 0;
        __syncthreads();
        sums[threadIdx.x] += v;
        __syncthreads();
    }
    unsigned at = sums[threadIdx.x] - n;
    // listCount[1]: a fingerprint of the list (which block at which position), for the host's cached copy of
    // what it derives from the list (spb_api.cu render_wavefront)
    unsigned h = 0;
    for (unsigned b = b0; b < b1; ++b)
        if (coverage[b])
        {
            h += (b + 1u) * (2654435761u * at + 1u);
            blockList[at++] = b;
        }
    if (h) atomicAdd(&listHash, h);
    __syncthreads();
    if (threadIdx.x == 1023) listCount[0] = sums[1023];
    if (threadIdx.x == 0) listCount[1] = listHash;
}

// Pixels whose block is not covered: the sample loop of sp_PathTraceTile
// (simd_path_tracer.cpp:216-321) with the miss branch only -- same functions, same order of
// operations as k_trace<PRIMARY> + k_shade_miss + k_accumulate, without the queues in between.
template <int MATH, int ENVFILTER>
__device__ __forceinline__ void sky_pixel_samples(const WaveArgs &a, const DMaterials &M, unsigned x, unsigned y,
                                                  Counters &cnt)
{
    const float weight = 1.0f / (float)a.spp;
    uint32_t pixelIndex = x + y * a.camera.width;
    f3 total = mk3(0.0f, 0.0f, 0.0f);
    for (unsigned s = 0; s < a.spp; ++s)
    {
        uint32_t rng = stream_seed(pixelIndex, s, a.frame);
        f3 o, d;
        primary_ray(a.camera, x, y, rng, o, d);
        f3 radiance = miss_radiance<MATH, ENVFILTER>(M, neg3(d), a.clampValue, &cnt);
        total = add3(total, mul3(radiance, weight));
    }
    v4f out;
    out.x = total.x; out.y = total.y; out.z = total.z; out.w = 1.0f;
    a.out[(size_t)pixelIndex] = out;
}

// One-lookup path: spb_core.cuh sky_one_lookup(); pixels it cannot settle (about 4 %) are listed for
// k_sky_listed, which runs the sample loop.
template <int MATH, int ENVFILTER>
__device__ __forceinline__ bool sky_pixel_one_lookup(const WaveArgs &a, const DMaterials &M, unsigned x, unsigned y)
{
    f3 total;
    if (!sky_one_lookup<MATH, ENVFILTER>(M, a.camera, x, y, a.skyDirectionSpread, a.spp, total)) return false;
    v4f out;
    out.x = total.x; out.y = total.y; out.z = total.z; out.w = 1.0f;
    a.out[(size_t)x + (size_t)y * a.camera.width] = out;
    return true;
}

template <int MATH, int ENVFILTER>
__global__ void __launch_bounds__(256)
k_sky(const __grid_constant__ WaveArgs a)
{
    const DMaterials &M = *a.materials;
    Counters cnt = {0, 0, 0, 0};
    unsigned shaded = 0;
    const unsigned stride = gridDim.x * blockDim.x;
    const unsigned rounds = (a.blocksX * a.blocksY * 32u + stride - 1) / stride;
    for (unsigned k = 0; k < rounds; ++k)
    {
        // 8 x 4 pixel blocks per warp, like the block mask
        unsigned p = k * stride + blockIdx.x * blockDim.x + threadIdx.x;
        unsigned blk = p >> 5, l = p & 31u;
        unsigned by = blk / a.blocksX, bx = blk - by * a.blocksX;
        unsigned x = a.x0 + bx * 8 + (l & 7u), y = a.y0 + by * 4 + (l >> 3);
        bool active = blk < a.blocksX * a.blocksY && x < a.x1 && y < a.y1 && a.blockMask[blk] == 0;
        bool listed = false;
        if (active)
        {
            if (a.skyList) listed = !sky_pixel_one_lookup<MATH, ENVFILTER>(a, M, x, y);
            else sky_pixel_samples<MATH, ENVFILTER>(a, M, x, y, cnt);
            shaded++;
        }
        if (a.skyList)
        {
            // warp-aggregated append of the pixels that need the sample loop
            unsigned mask = __ballot_sync(SPB_FULL, listed);
            if (mask)
            {
                unsigned base = 0;
                if (lane_id() == (unsigned)(__ffs(mask) - 1)) base = atomicAdd(&a.skyList[0], (unsigned)__popc(mask));
                base = __shfl_sync(SPB_FULL, base, __ffs(mask) - 1);
                if (listed) a.skyList[1 + base + __popc(mask & lanemask_lt())] = x | (y << 16);
            }
        }
        if (a.tileRowSky)
        {
            // the sky kernels' own cost class (their time is measured separately: spb_api.cu converts
            // units to nanoseconds class by class); a pixel settled by one lookup costs a sample's worth
            unsigned row = active ? (y / a.tileHeight - a.costRow0) : 0xFFFFFFFFu;
            unsigned peers = __match_any_sync(SPB_FULL, row);
            unsigned weight = (a.skyList && !listed) ? 1u : a.spp;
            unsigned total = __reduce_add_sync(peers, active ? weight : 0u);
            if (active && lane_id() == (unsigned)(__ffs(peers) - 1))
                atomicAdd(&a.tileRowSky[row], (unsigned long long)total * SPB_COST_SKY);
        }
    }
    shaded = __reduce_add_sync(SPB_FULL, shaded);
    if (lane_id() == 0 && shaded) atomicAdd(&a.stats[CTR_SKY_PIXELS], (unsigned long long)shaded);
    if (a.countStats)
    {
        unsigned e = __reduce_add_sync(SPB_FULL, cnt.envClamped);
        if (lane_id() == 0 && e) atomicAdd(&a.stats[CTR_ENV_CLAMPED], (unsigned long long)e);
    }
}

// the sky pixels k_sky could not settle with one lookup: the sample loop, every lane busy
template <int MATH, int ENVFILTER>
__global__ void __launch_bounds__(256)
k_sky_listed(const __grid_constant__ WaveArgs a)
{
    const DMaterials &M = *a.materials;
    Counters cnt = {0, 0, 0, 0};
    const unsigned count = a.skyList[0];
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x)
    {
        unsigned packed = a.skyList[1 + i];
        sky_pixel_samples<MATH, ENVFILTER>(a, M, packed & 0xFFFFu, packed >> 16, cnt);
    }
    if (a.countStats)
    {
        unsigned e = __reduce_add_sync(SPB_FULL, cnt.envClamped);
        if (lane_id() == 0 && e) atomicAdd(&a.stats[CTR_ENV_CLAMPED], (unsigned long long)e);
    }
}

// ---------------------------------------------------------------------------------------------
template <bool CULL, bool STATS, int MODE, bool SINGLE>
static void launch_trace_t(const WaveArgs &a, uint32_t bounce, unsigned grid, cudaStream_t stream)
{
    k_trace<CULL, STATS, MODE, SINGLE><<<grid, SPB_TRACE_THREADS, 0, stream>>>(a, bounce);
}

template <bool CULL, bool STATS, int MODE, bool SINGLE>
static unsigned trace_grid_t()
{
    static unsigned cached = 0;
    if (!cached)
    {
        int device = 0, sms = 0, perSm = 0;
        cudaGetDevice(&device);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_trace<CULL, STATS, MODE, SINGLE>, SPB_TRACE_THREADS, 0);
        if (perSm < 1) perSm = 1;
        cached = (unsigned)(sms * perSm); // persistent: every CTA resident, a multiple of the SM count
        // every warp in flight may leave one partly filled chunk in each queue: the slack the queues
        // and ray arrays carry (spb_api.cu render_wavefront) must cover them all
        if ((unsigned long long)cached * (SPB_TRACE_THREADS / 32) * SPB_CHUNK_MAX > SPB_QUEUE_SLACK)
        {
            fprintf(stderr, "[sp_b200] trace grid of %u CTAs needs more queue slack than SPB_QUEUE_SLACK provides\n", cached);
            abort();
        }
    }
    return cached;
}

#if defined(SPB_TRAV2)
#define SPB_TRACE_MODES(CALL, C, S)                                                \
    switch (mode) {                                                                \
    case SPB_TRACE_PRIMARY: CALL(C, S, SPB_TRACE_PRIMARY); break;                  \
    case SPB_TRACE_RESUME: break; /* (the A/B machine never evicts) */             \
    default: CALL(C, S, SPB_TRACE_QUEUE); break;                                   \
    }
#else
#define SPB_TRACE_MODES(CALL, C, S)                                                \
    switch (mode) {                                                                \
    case SPB_TRACE_PRIMARY: CALL(C, S, SPB_TRACE_PRIMARY); break;                  \
    case SPB_TRACE_EVICT: CALL(C, S, SPB_TRACE_EVICT); break;                      \
    case SPB_TRACE_RESUME: CALL(C, S, SPB_TRACE_RESUME); break;                    \
    default: CALL(C, S, SPB_TRACE_QUEUE); break;                                   \
    }
#endif
#define SPB_TRACE_DISPATCH(CALL)                                                   \
    do {                                                                           \
        int key = (cfg.cull ? 2 : 0) | (cfg.stats ? 1 : 0);                        \
        switch (key) {                                                             \
        case 0: SPB_TRACE_MODES(CALL, false, false); break;                        \
        case 1: SPB_TRACE_MODES(CALL, false, true); break;                         \
        case 2: SPB_TRACE_MODES(CALL, true, false); break;                         \
        default: SPB_TRACE_MODES(CALL, true, true); break;                         \
        }                                                                          \
    } while (0)

void launch_wave_trace(const KernelConfig &cfg, const WaveArgs &a, uint32_t bounce, int mode, cudaStream_t stream)
{
#if defined(SPB_TRAV2)
    if (mode == SPB_TRACE_RESUME) return;
#endif
    g_kernelLaunches++;
#if defined(SPB_TRAV2) || defined(SPB_NO_FUSED_ENTRY)
    const bool one = false;
#else
    const bool one = a.scene.objectCount == 1; // (C1-C4: the kernel without object entry / exit in its steps)
#endif
#define SPB_CALL(C, S, M) (one ? launch_trace_t<C, S, M, true>(a, bounce, trace_grid_t<C, S, M, true>(), stream) \
                               : launch_trace_t<C, S, M, false>(a, bounce, trace_grid_t<C, S, M, false>(), stream))
    SPB_TRACE_DISPATCH(SPB_CALL);
#undef SPB_CALL
}

static unsigned shade_grid()
{
    static unsigned cached = 0;
    if (!cached)
    {
        int device = 0, sms = 0;
        cudaGetDevice(&device);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        cached = (unsigned)sms * 8u;
    }
    return cached;
}

// Grid of a grid-stride shading kernel: every CTA resident at once (SM count x the kernel's own occupancy), so that
// all of them run for the whole launch.  With the fixed 8 CTAs per SM a kernel compiled to 40 registers (6 resident)
// ran 1.33 waves: the last third of its CTAs alone on the GPU, at a third of the memory-level parallelism a
// latency-bound kernel lives on.  (SPB_RESIDENT_SHADE_GRIDS 0: the fixed grid; 1: k_shade_miss only; 2: the hit kernels too.)
#ifndef SPB_RESIDENT_SHADE_GRIDS
#define SPB_RESIDENT_SHADE_GRIDS 2
#endif
template <class K>
static unsigned resident_grid(K kernel, int level)
{
    if (SPB_RESIDENT_SHADE_GRIDS < level) return shade_grid();
    // (per thread: a Library is driven by one thread, and every device of a multi-device host has its own)
    static thread_local std::map<const void *, unsigned> cache;
    unsigned &grid = cache[(const void *)kernel];
    if (!grid)
    {
        int device = 0, sms = 0, perSm = 0;
        cudaGetDevice(&device);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kernel, 256, 0);
        if (perSm < 1) perSm = 1;
        grid = (unsigned)(sms * perSm);
    }
    return grid;
}

#define SPB_SHADE_LAUNCH(MISS, HIT)                                                                   \
    do {                                                                                              \
        if (!a.fuseMiss) MISS<<<resident_grid(MISS, 1), 256, 0, stream>>>(a, bounce);                 \
        else g_kernelLaunches--; /* (the trace kernel shaded the escaped rays as it retired them) */  \
        HIT<<<resident_grid(HIT, 2), 256, 0, stream>>>(a, bounce);                                    \
    } while (0)

void launch_wave_shade(const KernelConfig &cfg, const WaveArgs &a, uint32_t bounce, cudaStream_t stream)
{
    g_kernelLaunches += 2;
    int key = (cfg.math ? 2 : 0) | (cfg.envFilter ? 1 : 0);
    switch (key)
    {
    case 0: SPB_SHADE_LAUNCH((k_shade_miss<0, 0>), (k_shade_hit<0, 0>)); break;
    case 1: SPB_SHADE_LAUNCH((k_shade_miss<0, 1>), (k_shade_hit<0, 1>)); break;
    case 2: SPB_SHADE_LAUNCH((k_shade_miss<1, 0>), (k_shade_hit<1, 0>)); break;
    default: SPB_SHADE_LAUNCH((k_shade_miss<1, 1>), (k_shade_hit<1, 1>)); break;
    }
}

void launch_wave_shade_sorted(const KernelConfig &cfg, const WaveArgs &a, uint32_t bounce, cudaStream_t stream)
{
    g_kernelLaunches += 2;
    int key = (cfg.math ? 2 : 0) | (cfg.envFilter ? 1 : 0);
    switch (key)
    {
    case 0: SPB_SHADE_LAUNCH((k_shade_miss<0, 0>), (k_shade_hit_tiles<0, 0>)); break;
    case 1: SPB_SHADE_LAUNCH((k_shade_miss<0, 1>), (k_shade_hit_tiles<0, 1>)); break;
    case 2: SPB_SHADE_LAUNCH((k_shade_miss<1, 0>), (k_shade_hit_tiles<1, 0>)); break;
    default: SPB_SHADE_LAUNCH((k_shade_miss<1, 1>), (k_shade_hit_tiles<1, 1>)); break;
    }
}

void launch_sky(const KernelConfig &cfg, const WaveArgs &a, cudaStream_t stream)
{
    g_kernelLaunches++;
    unsigned grid = shade_grid();
    int key = (cfg.math ? 2 : 0) | (cfg.envFilter ? 1 : 0);
    switch (key)
    {
    case 0: k_sky<0, 0><<<grid, 256, 0, stream>>>(a); break;
    case 1: k_sky<0, 1><<<grid, 256, 0, stream>>>(a); break;
    case 2: k_sky<1, 0><<<grid, 256, 0, stream>>>(a); break;
    default: k_sky<1, 1><<<grid, 256, 0, stream>>>(a); break;
    }
}

void launch_sky_listed(const KernelConfig &cfg, const WaveArgs &a, cudaStream_t stream)
{
    g_kernelLaunches++;
    unsigned grid = shade_grid();
    int key = (cfg.math ? 2 : 0) | (cfg.envFilter ? 1 : 0);
    switch (key)
    {
    case 0: k_sky_listed<0, 0><<<grid, 256, 0, stream>>>(a); break;
    case 1: k_sky_listed<0, 1><<<grid, 256, 0, stream>>>(a); break;
    case 2: k_sky_listed<1, 0><<<grid, 256, 0, stream>>>(a); break;
    default: k_sky_listed<1, 1><<<grid, 256, 0, stream>>>(a); break;
    }
}

void launch_coverage(const WaveArgs &a, uint64_t instancedTriangles, bool everything, uint8_t *coverage,
                     uint32_t *blockList, uint32_t *listCount, cudaStream_t stream)
{
    g_kernelLaunches += everything ? 1 : 2;
    const unsigned blocks = a.blocksX * a.blocksY;
    cudaMemsetAsync(coverage, 0, (size_t)blocks + 1, stream);
    if (everything) cudaMemsetAsync(coverage + blocks, 1, 1, stream);
    else if (instancedTriangles)
    {
        unsigned long long want = (instancedTriangles + 255) / 256;
        unsigned grid = (unsigned)(want < (unsigned long long)shade_grid() * 4 ? want : (unsigned long long)shade_grid() * 4);
        k_cover<<<grid, 256, 0, stream>>>(a, instancedTriangles, coverage);
    }
    k_list_blocks<<<1, 1024, 0, stream>>>(blocks, coverage, blockList, listCount);
}

void launch_candidates(const WaveArgs &a, uint32_t *candidates, cudaStream_t stream)
{
    g_kernelLaunches++;
    unsigned want = (a.bandBlocks * 32u + 127u) / 128u, grid = shade_grid() * 2u;
    if (want < grid) grid = want ? want : 1u;
    k_candidates<<<grid, 128, 0, stream>>>(a, candidates);
}

void launch_wave_accumulate(const WaveArgs &a, cudaStream_t stream)
{
    g_kernelLaunches++;
    k_accumulate<<<shade_grid(), 256, 0, stream>>>(a);
}

} // namespace spb
