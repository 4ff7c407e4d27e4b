// spb_render.cu -- the two integrator kernels (per-pixel streams and per-tile serial streams).
//
// Compiled four times, once per (SPB_INST_MATH, SPB_INST_ENV) pair, so the 16 template instances
// build in parallel (see __graft_entry__.build()).  Flags as for spb_kernels.cu: sm_100a,
// -fmad=false, -lineinfo.
#include "spb_kernels.cuh"

#ifndef SPB_INST_MATH
#error "compile with -DSPB_INST_MATH=0|1 -DSPB_INST_ENV=0|1"
#endif

namespace spb {

__device__ __forceinline__ unsigned warp_sum(unsigned v)
{
    return __reduce_add_sync(0xFFFFFFFFu, v);
}

// One thread per pixel; a warp covers an 8x4 pixel block (coherent primary rays), a CTA of 256
// threads covers 16x16.  Samples of a pixel are traced in order by the same thread so that
// sum_s radiance_s * (1/spp) is accumulated in the reference's order
// (simd_path_tracer.cpp:216-321).
template <int MATH, int ENVFILTER, bool CULL, bool STATS>
__global__ void __launch_bounds__(256)
k_render_pixels(RenderArgs a)
{
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned px = a.x0 + blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
    const unsigned py = a.y0 + blockIdx.y * 16 + (warp >> 1) * 4 + (lane >> 3);
    const bool active = px < a.x1 && py < a.y1;

    uint32_t stack[SPB_STACK_SIZE];
    float stackT[SPB_STACK_SIZE];
    PathCounters pc = {0, 0, 0};
    Counters ctr = {0, 0, 0, 0};
    unsigned paths = 0;

    if (active)
    {
        const DMaterials &M = *a.materials;
        const uint32_t pixelIndex = px + py * a.camera.width;
        const float weight = 1.0f / (float)a.spp;
        f3 total = mk3(0.0f, 0.0f, 0.0f);
        for (uint32_t s = 0; s < a.spp; ++s)
        {
            uint32_t rng = stream_seed(pixelIndex, s, a.frame);
            f3 radiance = trace_path<MATH, ENVFILTER, CULL>(a.scene, M, a.camera, px, py, rng,
                a.bounces, a.clampValue, stack, stackT, pc, STATS ? &ctr : nullptr);
            total = add3(total, mul3(radiance, weight));
            paths++;
        }
        v4f o;
        o.x = total.x; o.y = total.y; o.z = total.z; o.w = 1.0f;
        a.out[pixelIndex] = o;
    }

    // counters: one atomic per warp per slot
    unsigned sPaths = warp_sum(paths), sRays = warp_sum(pc.rays), sHits = warp_sum(pc.hits),
             sMiss = warp_sum(pc.misses);
    if (lane == 0)
    {
        atomicAdd(&a.counters[CTR_PATHS], (unsigned long long)sPaths);
        atomicAdd(&a.counters[CTR_RAYS], (unsigned long long)sRays);
        atomicAdd(&a.counters[CTR_HITS], (unsigned long long)sHits);
        atomicAdd(&a.counters[CTR_MISSES], (unsigned long long)sMiss);
        if (a.tileRowCost && sRays)
        {
            // every pixel row of a warp lies in one tile row when tileHeight % 4 == 0
            unsigned wy = a.y0 + blockIdx.y * 16 + (warp >> 1) * 4;
            unsigned row = (wy - (a.y0 / a.tileHeight) * a.tileHeight) / a.tileHeight;
            atomicAdd(&a.tileRowCost[row], (unsigned long long)sRays);
        }
    }
    if (STATS)
    {
        unsigned n = warp_sum(ctr.nodeVisits), t = warp_sum(ctr.triangleTests),
                 ob = warp_sum(ctr.objectTests), e = warp_sum(ctr.envClamped);
        if (lane == 0)
        {
            atomicAdd(&a.counters[CTR_NODE_VISITS], (unsigned long long)n);
            atomicAdd(&a.counters[CTR_TRIANGLE_TESTS], (unsigned long long)t);
            atomicAdd(&a.counters[CTR_OBJECT_TESTS], (unsigned long long)ob);
            atomicAdd(&a.counters[CTR_ENV_CLAMPED], (unsigned long long)e);
        }
    }
}

template <int MATH, int ENVFILTER, bool CULL, bool STATS>
static void launch_render_t(const RenderArgs &args, cudaStream_t stream)
{
    dim3 grid((args.x1 - args.x0 + 15) / 16, (args.y1 - args.y0 + 15) / 16);
    k_render_pixels<MATH, ENVFILTER, CULL, STATS><<<grid, 256, 0, stream>>>(args);
}

#define SPB_JOIN_(a, b, c, d) a##b##c##d
#define SPB_JOIN(a, b, c, d) SPB_JOIN_(a, b, c, d)
#define SPB_RENDER_NAME SPB_JOIN(launch_render_m, SPB_INST_MATH, e, SPB_INST_ENV)
#define SPB_TILES_NAME SPB_JOIN(launch_tiles_m, SPB_INST_MATH, e, SPB_INST_ENV)

void SPB_RENDER_NAME(const KernelConfig &cfg, const RenderArgs &args, cudaStream_t stream)
{
    if (cfg.cull)
    {
        if (cfg.stats) launch_render_t<SPB_INST_MATH, SPB_INST_ENV, true, true>(args, stream);
        else launch_render_t<SPB_INST_MATH, SPB_INST_ENV, true, false>(args, stream);
    }
    else
    {
        if (cfg.stats) launch_render_t<SPB_INST_MATH, SPB_INST_ENV, false, true>(args, stream);
        else launch_render_t<SPB_INST_MATH, SPB_INST_ENV, false, false>(args, stream);
    }
}

// ---------------------------------------------------------------------------------------------
// sp_PathTraceTile: the rng stream is consumed serially across the pixels of a tile and the
// number of draws per sample depends on the path, so a tile is inherently one thread of control.
template <int MATH, int ENVFILTER, bool CULL, bool STATS>
__global__ void k_tiles_serial(TileArgs a)
{
    unsigned tile = blockIdx.x * blockDim.x + threadIdx.x;
    if (tile >= a.count) return;
    uint32_t stack[SPB_STACK_SIZE];
    float stackT[SPB_STACK_SIZE];
    const DMaterials &M = *a.materials;
    uint32_t minX = a.tiles[tile * 4 + 0], minY = a.tiles[tile * 4 + 1];
    uint32_t maxX = a.tiles[tile * 4 + 2], maxY = a.tiles[tile * 4 + 3];
    uint32_t rng = a.rngStates[tile];
    PathCounters pc = {0, 0, 0};
    Counters ctr = {0, 0, 0, 0};
    unsigned long long paths = 0;
    long long start = clock64();
    const float weight = 1.0f / (float)a.spp;
    for (uint32_t y = minY; y < maxY; ++y)
    {
        for (uint32_t x = minX; x < maxX; ++x)
        {
            f3 total = mk3(0.0f, 0.0f, 0.0f);
            for (uint32_t s = 0; s < a.spp; ++s)
            {
                f3 radiance = trace_path<MATH, ENVFILTER, CULL>(a.scene, M, a.camera, x, y, rng,
                    a.bounces, a.clampValue, stack, stackT, pc, STATS ? &ctr : nullptr);
                total = add3(total, mul3(radiance, weight));
                paths++;
            }
            v4f o;
            o.x = total.x; o.y = total.y; o.z = total.z; o.w = 1.0f;
            a.out[x + y * a.camera.width] = o;
        }
    }
    a.rngStates[tile] = rng;
    unsigned long long *c = a.counters + (size_t)tile * CTR_COUNT;
    c[CTR_PATHS] = paths;
    c[CTR_RAYS] = pc.rays;
    c[CTR_HITS] = pc.hits;
    c[CTR_MISSES] = pc.misses;
    c[CTR_NODE_VISITS] = ctr.nodeVisits;
    c[CTR_TRIANGLE_TESTS] = ctr.triangleTests;
    c[CTR_OBJECT_TESTS] = ctr.objectTests;
    c[CTR_ENV_CLAMPED] = ctr.envClamped;
    c[CTR_CLOCK_SUM] = (unsigned long long)(clock64() - start);
}

template <int MATH, int ENVFILTER, bool CULL, bool STATS>
static void launch_tiles_t(const TileArgs &args, cudaStream_t stream)
{
    // one tile per thread, spread one thread per CTA over the SMs first (tiles are long serial
    // jobs; co-locating them in a warp would only serialise their divergent paths)
    unsigned threads = args.count > 148 * 8 ? 32 : 1;
    unsigned blocks = (args.count + threads - 1) / threads;
    k_tiles_serial<MATH, ENVFILTER, CULL, STATS><<<blocks, threads, 0, stream>>>(args);
}

void SPB_TILES_NAME(const KernelConfig &cfg, const TileArgs &args, cudaStream_t stream)
{
    if (cfg.cull)
    {
        if (cfg.stats) launch_tiles_t<SPB_INST_MATH, SPB_INST_ENV, true, true>(args, stream);
        else launch_tiles_t<SPB_INST_MATH, SPB_INST_ENV, true, false>(args, stream);
    }
    else
    {
        if (cfg.stats) launch_tiles_t<SPB_INST_MATH, SPB_INST_ENV, false, true>(args, stream);
        else launch_tiles_t<SPB_INST_MATH, SPB_INST_ENV, false, false>(args, stream);
    }
}


} // namespace spb
