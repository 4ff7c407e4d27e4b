// spb_kernels.cuh -- launch interface between the C ABI (spb_api.cu) and the kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "spb_bvh.h"
#include "spb_core.cuh"

namespace spb {

// slots of the device counter block every kernel accumulates into
enum
{
    CTR_PATHS = 0, CTR_RAYS, CTR_HITS, CTR_MISSES, CTR_NODE_VISITS, CTR_TRIANGLE_TESTS,
    CTR_OBJECT_TESTS, CTR_ENV_CLAMPED, CTR_CLOCK_SUM, CTR_SKY_PIXELS, CTR_COUNT
};

struct KernelConfig
{
    int math;      // SP_B200_MATH_*
    int envFilter; // SP_B200_ENV_*
    int cull;      // 0 / 1
    int stats;     // 0 / 1: count node visits and triangle tests
};

struct RenderArgs
{
    DScene scene;
    const DMaterials *materials;
    DCamera camera;
    uint32_t x0, y0, x1, y1; // pixel rectangle
    uint32_t spp, bounces, frame;
    float clampValue;
    v4f *out;                         // full image, indexed x + y * camera.width
    unsigned long long *counters;     // CTR_COUNT slots
    unsigned long long *tileRowCost;  // may be null; rays per tile row, row 0 = y0 / tileHeight
    uint32_t tileHeight;
};

// number of kernels this library has launched since load (bench.py's gpu_launches)
extern std::atomic<unsigned long long> g_kernelLaunches; // (several host threads launch in multi-device mode)

void launch_render(const KernelConfig &cfg, const RenderArgs &args, cudaStream_t stream);

// sp_PathTraceTile semantics: one serial XorShift32 stream per tile (simd_path_tracer.cpp:178-345)
struct TileArgs
{
    DScene scene;
    const DMaterials *materials;
    DCamera camera;
    const uint32_t *tiles;  // count x 4: minX minY maxX maxY (already clamped to the image)
    uint32_t *rngStates;    // count, in/out
    uint32_t count;
    uint32_t spp, bounces;
    float clampValue;
    v4f *out;
    unsigned long long *counters; // count x CTR_COUNT (per tile)
};
void launch_tiles_serial(const KernelConfig &cfg, const TileArgs &args, cudaStream_t stream);

struct HitRecord
{
    float t;
    uint32_t materialId;
    float nx, ny, nz;
    float u, v;
    int32_t triangle;
    int32_t object;
};

void launch_primary_hits(const KernelConfig &cfg, const DScene &scene, const DCamera &camera,
                         uint32_t sample, uint32_t frame, int32_t *tri, int32_t *obj, float *t,
                         unsigned long long *counters, cudaStream_t stream);
void launch_intersect_batch(const KernelConfig &cfg, const DScene &scene, uint32_t count,
                            const float *origins3, const float *dirs3, HitRecord *out,
                            unsigned long long *counters, cudaStream_t stream);
// sp_RayIntersectMesh: object-space ray against object 0's mesh; normal left in object space
void launch_intersect_mesh(const KernelConfig &cfg, const DScene &scene, uint32_t smooth,
                           const float *origin3, const float *dir3, HitRecord *out,
                           cudaStream_t stream);
// bvh_IntersectRay semantics on object 0's mesh: every leaf whose own box the ray passes
void launch_collect_leaves(const DScene &scene, const float *origin3, const float *dir3,
                           uint32_t *leaves, uint32_t maxLeaves, uint32_t *countAndError,
                           cudaStream_t stream);
// running mean over frames: accum += (frame - accum) / (framesAccumulated + 1)
void launch_accumulate_frame(v4f *accum, const v4f *frame, uint32_t count, uint32_t framesAccumulated, cudaStream_t stream);
// batched forms (the reference's perf tests): per ray leaf count / xor / sum; per ray t and triangle index
void launch_collect_leaves_batch(const DScene &scene, uint32_t count, const float *origins3, const float *dirs3,
                                 uint32_t *out3, cudaStream_t stream);
void launch_intersect_mesh_batch(const KernelConfig &cfg, const DScene &scene, uint32_t count, const float *origins3,
                                 const float *dirs3, float *tOut, int32_t *triOut, cudaStream_t stream);
// the device forms of simd_RayIntersectAabb4 on caller data (k_slab_kat): known-answer tests
void launch_slab_kat(uint32_t count, const float *boxMin12, const float *boxMax12, const float *origin3, const float *invDir3,
                     uint32_t *masks, float *tnear, cudaStream_t stream);
// ComputeRadianceForPath over n vertices of 15 floats (materialId bits, P3, out3, in3, n3, uv2)
void launch_radiance_for_path(const KernelConfig &cfg, const DMaterials *materials,
                              const float *path15, uint32_t n, float clampValue, float *out3,
                              cudaStream_t stream);
// sp_EvaluateMaterial with an explicit material (not looked up by id); out7 = albedo3 emission3 r
void launch_evaluate_material(const KernelConfig &cfg, const DMaterials *materials,
                              uint32_t materialSlot, const float *vertex15, float *out7,
                              cudaStream_t stream);

// Output stage: PerformToneMapping (post_processing.frag.glsl:19-26) on RGBA f32 pixels and the
// 8-bit UNORM store; out = count x RGBA8
void launch_tone_map(const KernelConfig &cfg, const v4f *pixels, uint32_t count, float exposure, uint32_t *out,
                     cudaStream_t stream);

// Environment pre-processing (spb_cubemap.cu; src/cubemap.cpp:108-291).  out = 6 faces of
// width x height RGBA f32, layer-major (+X -X +Y -Y +Z -Z), rows top to bottom.
void launch_cube_map(const KernelConfig &cfg, const DImage &env, v4f *out, uint32_t width, uint32_t height,
                     cudaStream_t stream);
struct IrradianceArgs
{
    DImage env;
    v4f *out;
    uint32_t width, height;
    float clampValue;
    // uniform mode: the phi / theta values of the reference's float-accumulating loops
    const float *phis, *thetas;
    uint32_t phiCount, thetaCount;
    // random mode: serial XorShift32 stream, three draws per sample; jump tables (32 x 32 words,
    // see xorshift_jump) for strides of 3 * samplesPerPixel draws (texel) and 3 draws (sample)
    uint32_t samplesPerPixel, seed;
    float sampleContribution;
    const uint32_t *jumpTexel, *jumpSample;
};
void launch_irradiance(const KernelConfig &cfg, const IrradianceArgs &args, int mode, cudaStream_t stream);

// Device BVH builder (spb_lbvh.cu): Morton keys, radix sort, binary radix tree, bottom-up boxes.
// false on a CUDA error; the tree is validated and collapsed by bvh4_from_binary (spb_bvh.h).
bool lbvh_build_binary_device(const float *aabbMin, const float *aabbMax, uint32_t count, BinaryTree *tree,
                              float *kernelMs, cudaStream_t stream);
// The whole build on the device, 4-wide collapse included (spb_lbvh.cu): the finished tree comes back and
// bvh4_adopt_device_tree (spb_bvh.cpp) checks it.  false on a CUDA error or a tree the emission cannot finish.
bool lbvh_build_bvh4_device(const float *aabbMin, const float *aabbMax, uint32_t count, DeviceTree4 *tree,
                            float *kernelMs, cudaStream_t stream);

// ---------------------------------------------------------------------------------------------
// wavefront renderer (spb_wavefront.cu)

#ifndef SPB_TRACE_THREADS
#define SPB_TRACE_THREADS 128
#endif
#ifndef SPB_TRACE_MIN_BLOCKS
#define SPB_TRACE_MIN_BLOCKS 5
#endif
// A/B knob: > 0 stages that many top-level TLAS nodes in shared memory (spb_wavefront.cu k_trace)
#ifndef SPB_TLAS_SMEM_NODES
#define SPB_TLAS_SMEM_NODES 0
#endif
// tile-row cost units (sp_b200_RenderRows tileRowCost): per sky-kernel sample / escaped ray / surface hit
// (measured on C3: a sky-kernel sample ~5 ps since most sky pixels take one lookup, an escaped ray
// through the queues ~85 ps, a surface hit with everything it triggers ~400 ps)
#define SPB_COST_SKY 1u
#define SPB_COST_MISS 16u
#define SPB_COST_HIT 80u
// a warp keeps walking until fewer than this many of its lanes still have a node to visit,
// then retires the finished lanes and refills them from the queue
#ifndef SPB_REFILL_THRESHOLD
#define SPB_REFILL_THRESHOLD 12u
#endif

// per-(pass, bounce) device counters
// RAYS: length of this bounce's ray queue (slots handed to the trace kernel; may contain holes);
// HITS / MISSES: ALLOCATED lengths of the hit / miss queues -- warps reserve chunks of slots, so a
// queue ends with up to one partly filled chunk per warp whose tail is holes (SPB_QUEUE_HOLE);
// CURSOR: work hand-out position of the trace kernel; NHITS / NMISSES: exact counts.
// CONT: continuation records written by an EVICT launch of the trace kernel (may exceed the buffer's capacity:
// the excess was not written); CONT_CURSOR: hand-out position of the RESUME launch.
enum { WCTR_RAYS = 0, WCTR_HITS, WCTR_MISSES, WCTR_CURSOR, WCTR_NHITS, WCTR_NMISSES, WCTR_CONT, WCTR_CONT_CURSOR, WCTR_STRIDE };
#define SPB_QUEUE_HOLE 0xFFFFFFFFu
// largest chunk a warp reserves from a queue counter with one atomic
#ifndef SPB_CHUNK_MAX
#define SPB_CHUNK_MAX 64u
#endif
// extra slots every queue / ray array carries for partly filled chunks: warps in flight x chunk
#define SPB_QUEUE_SLACK (8192u * SPB_CHUNK_MAX)
// modes of the trace kernel (spb_wavefront.cu k_trace)
enum { SPB_TRACE_QUEUE = 0, SPB_TRACE_PRIMARY = 1, SPB_TRACE_RESUME = 2, SPB_TRACE_EVICT = 3 };
// continuation record of an evicted straggler: 3 quads of state + SPB_CONT_STACK stack entries, two per quad
#define SPB_CONT_STACK 10u
#define SPB_CONT_QUADS (3u + SPB_CONT_STACK / 2u)

struct WaveArgs
{
    DScene scene;
    const DMaterials *materials;
    DCamera camera;
    uint32_t x0, y0, x1, y1;   // pixel rectangle of the strip being rendered
    uint32_t blocksX, blocksY; // 8x4 pixel blocks of the strip
    // Blocks that some triangle may project into (coverage pass), in row-major order; the pass at
    // hand renders bandBlocks of them starting at blockList[0].  Pixels of the other blocks are
    // the sky kernel's.
    const uint32_t *blockList;
    const uint8_t *blockMask;  // per block of the strip: 1 = in the list
    uint32_t bandBlocks;
    uint32_t workItems;        // bandBlocks * 32 * samplesThisPass; item = (block, pixel in block, sample)
    uint32_t samplesThisPass, firstSample, spp, bounces, frame;
    uint32_t pathCapacity;     // paths per pass the per-path arrays are sized for
    float clampValue;
    v4f *rays[2];              // ray records, two float4 each: (o, rng bits) (d, path id)
    v4f *hitRec;               // per ray slot: t, triangle slot bits, object index bits, -
    uint32_t *hitQ, *missQ;    // compact lists of ray slots
    v4f *pathTerms;            // [bounce][path] two float4: (E, cosine) (W, -)
    v4f *rad;                  // [path]: block-major, pixel in block, samplesThisPass per pixel
    v4f *out;                  // full image
    uint32_t *ctr;             // this pass's counters: bounces x WCTR_STRIDE
    unsigned long long *stats; // CTR_* slots (always valid; node/triangle counts only in stats launches)
    unsigned long long *tileRowCost; // may be null: cost units of the queue kernels per tile row (SPB_COST_MISS / _HIT)
    unsigned long long *tileRowSky;  // may be null: cost units of the sky kernels per tile row (SPB_COST_SKY per sample run)
    uint32_t tileHeight;
    uint32_t costRow0;         // tile row that tileRowCost[0] stands for
    int countStats;            // 1: stats launch
    // Primary hits shaded tile by tile (SPB_SORT_TILE items) with the bounce rays of a tile written
    // in direction order: the primary trace kernel leaves its results in hitRec by item instead of
    // a hit queue, `stage` holds the unsorted rays of a tile between the two phases.
    int sortPrimaryHits;
    v4f *stage;
    // a warp of the trace kernel retires and refills its lanes when fewer than this many are still
    // walking: 1 = packet mode (the whole warp starts and ends together; best when its rays are
    // coherent), SPB_REFILL_THRESHOLD otherwise.  Set per launch by the host.
    uint32_t refillThreshold;
    // EVICT / RESUME launches: the continuation buffer (SPB_CONT_QUADS quads per record) and its capacity in records
    v4u *cont;
    uint32_t contCapacity;
    // Per pixel of the pass (block-major, like the items): candidate triangles of its camera rays
    // (k_candidates): SPB_CAND_STRIDE words per pixel, [0] = count or SPB_CAND_FALLBACK, then the
    // triangle slots.  Null: every primary ray walks the tree.
    const uint32_t *candidates;
    // k_sky: > 0 enables the one-lookup path for pixels whose samples provably read one texel;
    // the value bounds |direction of any sample - direction of the pixel centre| (host: jitter
    // and pixel angle of the camera).  skyList[0] counts, skyList[1..] lists the other pixels.
    float skyDirectionSpread;
    uint32_t *skyList;
    // 0: escaped rays go to the miss queue and k_shade_miss shades them; else 1 | mathMode << 1 | envFilter << 2:
    // the trace kernel shades a ray that escapes where it retires it (no queue entry, no k_shade_miss launch)
    uint32_t fuseMiss;
};
// (SPB_CAND_MAX / SPB_CAND_STRIDE / SPB_CAND_FALLBACK: spb_core.cuh)
// items (pixel x sample, block-major) whose bounce rays are ordered together
#ifndef SPB_SORT_TILE
#define SPB_SORT_TILE 2048u
#endif

// Coverage pass: marks every 8x4 pixel block of the strip that the (2-pixel padded) screen bounding
// box of some triangle touches; coverage[blocks] = 1 flags "everything" (a triangle straddles the
// camera plane).  Then the marked blocks are listed in row-major order; listCount[0] = how many.
void launch_coverage(const WaveArgs &args, uint64_t instancedTriangles, bool everything, uint8_t *coverage,
                     uint32_t *blockList, uint32_t *listCount, cudaStream_t stream);
void launch_wave_trace(const KernelConfig &cfg, const WaveArgs &args, uint32_t bounce, int mode, cudaStream_t stream);
void launch_wave_shade(const KernelConfig &cfg, const WaveArgs &args, uint32_t bounce, cudaStream_t stream);
// k_shade_miss + the tile kernel, which writes the next bounce's rays in direction order per tile
// (bounce 0 needs sortPrimaryHits: primary results by item; not for the last bounce)
void launch_wave_shade_sorted(const KernelConfig &cfg, const WaveArgs &args, uint32_t bounce, cudaStream_t stream);
void launch_wave_accumulate(const WaveArgs &args, cudaStream_t stream);
// single-object scenes: candidate triangles per pixel of the pass (see k_candidates)
void launch_candidates(const WaveArgs &args, uint32_t *candidates, cudaStream_t stream);
// Pixels of the strip whose block is not in the list: every sample's camera ray leaves the scene
// untouched, so the whole pixel is evaluated in one thread: ray generation, background material,
// accumulation in sample order.  Adds the pixels it shaded to stats[CTR_SKY_PIXELS] and their cost
// to tileRowCost.
void launch_sky(const KernelConfig &cfg, const WaveArgs &args, cudaStream_t stream);
// the pixels k_sky listed in skyList (full sample loop); reads the count from the device
void launch_sky_listed(const KernelConfig &cfg, const WaveArgs &args, cudaStream_t stream);

} // namespace spb
