/* sp_b200.h -- C ABI of libspb200.so, the B200 (sm_100a) implementation of vk_cinematic's
 * CPU "SIMD" path tracer (the sp_ path).
 *
 * The entry points keep the reference's names, argument meaning and error behaviour so that the
 * library drops in where the reference unity-includes sp_scene.cpp, sp_material_system.cpp and
 * simd_path_tracer.cpp (reference src/main.cpp:235-240).  Each declaration cites the reference
 * interface it replaces as `file:line` relative to the reference checkout.  The POD types below
 * are byte-for-byte layout-compatible with the reference's (sizes asserted at the bottom; see
 * SURVEY.md §8 for the probed values).  A reference-side translation unit that already includes
 * the reference's own headers defines SP_B200_USE_REFERENCE_TYPES before including this file
 * (see INTEGRATION.md).
 *
 * Everything computes on the GPU.  There is no CPU fallback: every compute entry point aborts
 * through the log callback (Assert convention, reference src/platform.h:18-22) when no CUDA
 * device is usable.
 */
#ifndef SP_B200_H
#define SP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef SP_B200_USE_REFERENCE_TYPES

typedef uint8_t u8;
typedef int32_t i32;
typedef uint32_t u32;
typedef uint64_t u64;
typedef float f32;
typedef double f64;
typedef u32 b32;

/* math_lib.h:12-92 (unions there; plain structs with the same layout here) */
typedef struct vec2 { f32 x, y; } vec2;
typedef struct vec3 { f32 x, y, z; } vec3;
typedef struct vec4 { f32 x, y, z, w; } vec4;
typedef struct mat4 { vec4 columns[4]; } mat4;
typedef vec4 quat; /* (x, y, z, w) */

/* platform.h:87-92 */
typedef struct MemoryArena { void *base; u64 size; u64 capacity; } MemoryArena;
/* memory_pool.h:3-9 */
typedef struct MemoryPool { u8 *storage; u32 objectSize; u32 capacity; u32 headIndex; } MemoryPool;
/* bvh.h:14-18.  `root` holds this library's acceleration-structure handle, not a bvh_Node. */
typedef struct bvh_Tree { void *root; MemoryPool memoryPool; } bvh_Tree;
/* aabb.h:3-7 */
typedef struct Aabb { vec3 min; vec3 max; } Aabb;
/* mesh.h:11-16 */
typedef struct VertexPNT { vec3 position; vec3 normal; vec2 textureCoord; } VertexPNT;
/* asset_loader/asset_loader.h:13-18 */
typedef struct HdrImage { float *pixels; uint32_t width; uint32_t height; } HdrImage;
/* tile.h:3-9 */
typedef struct Tile { u32 minX, minY, maxX, maxY; } Tile;
/* math_utils.h:184-187 */
typedef struct RandomNumberGenerator { u32 state; } RandomNumberGenerator;
/* work_queue.h:4-11 */
typedef struct WorkQueue {
    volatile i32 head; volatile i32 tail; u32 objectSize; u32 maxObjects; void *buffer;
} WorkQueue;

/* sp_metrics.h:3-48 */
enum {
    sp_Metric_CyclesElapsed,
    sp_Metric_PathsTraced,
    sp_Metric_RaysTraced,
    sp_Metric_RayHitCount,
    sp_Metric_RayMissCount,
    sp_Metric_CyclesElapsed_RayIntersectScene,
    sp_Metric_CyclesElapsed_RayIntersectBroadphase,
    sp_Metric_CyclesElapsed_RayIntersectMesh,
    sp_Metric_CyclesElapsed_RayIntersectMeshMidphase,
    sp_Metric_CyclesElapsed_RayIntersectTriangle,
    sp_Metric_RayIntersectMesh_MidphaseAabbTestCount,
    sp_Metric_RayIntersectMesh_TestsPerformed,
    SP_MAX_METRICS
};
typedef struct sp_Metrics { u64 values[SP_MAX_METRICS]; } sp_Metrics;

/* sp_scene.h:3-29 */
typedef struct sp_Mesh {
    VertexPNT *vertices;
    u32 *indices;
    u32 vertexCount;
    u32 indexCount;
    bvh_Tree midphaseTree;
    b32 useSmoothShading;
} sp_Mesh;

#define SP_SCENE_MAX_OBJECTS 32
typedef struct sp_Scene {
    vec3 aabbMin[SP_SCENE_MAX_OBJECTS];
    vec3 aabbMax[SP_SCENE_MAX_OBJECTS];
    sp_Mesh meshes[SP_SCENE_MAX_OBJECTS];
    u32 materials[SP_SCENE_MAX_OBJECTS];
    mat4 invModelMatrices[SP_SCENE_MAX_OBJECTS];
    mat4 modelMatrices[SP_SCENE_MAX_OBJECTS];
    u32 objectCount;
    MemoryArena memoryArena;
    bvh_Tree broadphaseTree;
} sp_Scene;

/* ray_intersection.h:3-8 */
typedef struct RayIntersectTriangleResult { f32 t; vec2 uv; vec3 normal; } RayIntersectTriangleResult;
/* sp_scene.h:31-37 */
typedef struct sp_RayIntersectMeshResult {
    RayIntersectTriangleResult triangleIntersection;
} sp_RayIntersectMeshResult;
/* sp_scene.h:39-53 */
typedef struct sp_RayIntersectSceneResult { f32 t; u32 materialId; vec3 normal; vec2 uv; } sp_RayIntersectSceneResult;

/* sp_material_system.h:3-46 */
typedef struct sp_Material {
    vec3 albedo; u32 albedoTexture; vec3 emission; u32 emissionTexture; f32 roughness;
} sp_Material;
typedef struct sp_MaterialOutput { vec3 albedo; vec3 emission; f32 roughness; } sp_MaterialOutput;
typedef struct sp_PathVertex {
    u32 materialId; vec3 worldPosition; vec3 outgoingDir; vec3 incomingDir; vec3 normal; vec2 uv;
} sp_PathVertex;
#define SP_MAX_MATERIALS 32
#define SP_MAX_IMAGES 16
typedef struct sp_MaterialSystem {
    u32 keys[SP_MAX_MATERIALS];
    sp_Material materials[SP_MAX_MATERIALS];
    u32 count;
    u32 imageKeys[SP_MAX_IMAGES];
    HdrImage images[SP_MAX_IMAGES];
    u32 imageCount;
    u32 backgroundMaterialId;
} sp_MaterialSystem;

/* simd_path_tracer.h:3-46 */
typedef struct ImagePlane { vec4 *pixels; u32 width; u32 height; } ImagePlane;
typedef struct Basis { vec3 right; vec3 up; vec3 forward; } Basis;
typedef struct sp_Camera {
    Basis basis; vec3 position; vec3 filmCenter; ImagePlane *imagePlane;
    f32 halfPixelWidth; f32 halfPixelHeight; f32 halfFilmWidth; f32 halfFilmHeight;
} sp_Camera;
typedef struct sp_Context { sp_Camera *camera; sp_Scene *scene; sp_MaterialSystem *materialSystem; } sp_Context;

#endif /* SP_B200_USE_REFERENCE_TYPES */

/* ============================ the reference's sp_ surface ============================ */

/* sp_scene.cpp:1-5.  The arena is recorded, never allocated from: acceleration structures live
 * in library-owned host and device memory. */
void sp_InitializeScene(sp_Scene *scene, MemoryArena *arena);
/* sp_scene.cpp:8-19.  Aliases caller memory exactly like the reference (C has no default
 * arguments: pass useSmoothShading explicitly). */
sp_Mesh sp_CreateMesh(VertexPNT *vertices, u32 vertexCount, u32 *indices, u32 indexCount,
                      b32 useSmoothShading);
/* sp_scene.cpp:21-54.  Snapshots vertices/indices and builds the 4-wide midphase BVH
 * (replaces bvh_CreateTree, bvh.cpp:51-200).  Arenas unused.  Asserts indexCount % 3 == 0. */
void sp_BuildMeshMidphase(sp_Mesh *mesh, MemoryArena *arena, MemoryArena *tempArena);
/* sp_scene.cpp:77-117.  Asserts objectCount < SP_SCENE_MAX_OBJECTS like the reference (use
 * sp_b200_AddObjectToScene for larger scenes). */
void sp_AddObjectToScene(sp_Scene *scene, sp_Mesh mesh, u32 material, vec3 position,
                         quat orientation, vec3 scale);
/* sp_AddObjectToScene without the 32-object cap: objects beyond the fixed table of sp_Scene are
 * kept by the library (per sp_Scene) and joined in by sp_BuildSceneBroadphase.  Returns the
 * object's index.  sp_b200_SceneObjectCount = table + kept objects. */
u32 sp_b200_AddObjectToScene(sp_Scene *scene, sp_Mesh mesh, u32 material, vec3 position,
                             quat orientation, vec3 scale);
u32 sp_b200_SceneObjectCount(sp_Scene *scene);
/* sp_scene.cpp:119-125.  Builds the object-level BVH and uploads the whole scene to the current
 * device once; later render calls reuse it. */
void sp_BuildSceneBroadphase(sp_Scene *scene);
/* sp_scene.cpp:229-339 / 127-227: single-ray forms (one-ray GPU batches; tests use them). */
sp_RayIntersectSceneResult sp_RayIntersectScene(sp_Scene *scene, vec3 rayOrigin,
                                                vec3 rayDirection, sp_Metrics *metrics);
sp_RayIntersectMeshResult sp_RayIntersectMesh(sp_Mesh mesh, vec3 rayOrigin, vec3 rayDirection,
                                              sp_Metrics *metrics);

/* sp_material_system.cpp:1-13, 15-29, 31-43, 45-57, 59-105 */
b32 sp_RegisterMaterial(sp_MaterialSystem *materialSystem, sp_Material material, u32 id);
sp_Material *sp_FindMaterialById(sp_MaterialSystem *materialSystem, u32 id);
HdrImage *sp_FindTexture(sp_MaterialSystem *materialSystem, u32 id);
b32 sp_RegisterTexture(sp_MaterialSystem *materialSystem, HdrImage image, u32 id);
sp_MaterialOutput sp_EvaluateMaterial(sp_MaterialSystem *materialSystem, sp_Material *material,
                                      sp_PathVertex *vertex);

/* simd_path_tracer.cpp:1-36, 38-63 (pure host arithmetic, same operation order) */
void sp_ConfigureCamera(sp_Camera *camera, ImagePlane *imagePlane, vec3 position, quat rotation,
                        f32 filmDistance);
u32 sp_CalculateFilmPositions(sp_Camera *camera, vec3 *filmPositions, vec2 *pixelPositions,
                              u32 count);
/* simd_path_tracer.cpp:107-175 (evaluated on the GPU by the same device code the renderer uses) */
vec3 ComputeRadianceForPath(sp_PathVertex *path, u32 pathLength, sp_MaterialSystem *materialSystem);
/* simd_path_tracer.cpp:178-345: the tile-render entry.  Serial XorShift32 stream over the tile's
 * pixels exactly as the reference (one GPU thread per tile: drop-in fidelity, not speed);
 * overwrites metrics[CyclesElapsed] with device nanoseconds, increments the counters. */
void sp_PathTraceTile(sp_Context *ctx, Tile tile, RandomNumberGenerator *rng, sp_Metrics *metrics);
/* sp_PathTraceTile may be called from many host threads at once, like WorkerThread does
 * (main.cpp:728-759): concurrent calls are combined -- every request that queues up while a launch is
 * running goes into the next launch, one GPU thread per tile -- instead of taking turns on a mutex.
 * Counts so far: launches made and tiles rendered by them. */
void sp_b200_TileCombinerStats(u64 *launches, u64 *tiles);

/* The reference DEFINES the next five `inline` in its headers (aabb.h:29-58, tile.h:11-42,
 * work_queue.h:13-44), so a reference-side translation unit already has them -- with C++ linkage --
 * and keeps using them; the library exports C versions of the same arithmetic for every other host
 * (tests/dropin compiles the reference-side case, tests/test_abi.py the plain C one). */
#ifndef SP_B200_USE_REFERENCE_TYPES
Aabb TransformAabb(vec3 boxMin, vec3 boxMax, vec3 position, quat orientation, vec3 scale);
u32 ComputeTiles(u32 totalWidth, u32 totalHeight, u32 tileWidth, u32 tileHeight, Tile *tiles,
                 u32 maxTiles);
WorkQueue CreateWorkQueue(MemoryArena *arena, u32 objectSize, u32 maxObjects);
b32 WorkQueuePush(WorkQueue *queue, void *object, u32 objectSize);
void *WorkQueuePop(WorkQueue *queue, u32 objectSize);
#endif

/* The tile scheduler on top of the queue (main.cpp:246-250 sp_Task; :819-844 AddRayTracingWorkQueue;
 * :728-759 WorkerThread + g_metricsBuffer).  `internal` in the reference, so they carry the library
 * prefix here.  sp_b200_AddRayTracingWorkQueue: asserts the queue is empty, resets tail/head and
 * pushes one sp_Task per tile of ctx's image plane (tile size from sp_b200_Params, default 64x64 =
 * TILE_WIDTH/HEIGHT; at most queue->maxObjects tiles where the reference has MAX_TILES); returns the
 * tile count.  sp_b200_DrainRayTracingWorkQueue: what the worker threads do until the queue is empty
 * -- pop a task, RandomNumberGenerator{0xF51C0E49}, sp_PathTraceTile, append the tile's sp_Metrics --
 * with every popped tile of a context rendered in one launch; pixels equal the reference's, metrics
 * are appended in pop order (at most maxMetrics; metricsBuffer may be NULL); returns tasks rendered. */
#ifndef SP_B200_USE_REFERENCE_TYPES /* main.cpp defines its own sp_Task (same layout, 24 bytes) */
typedef struct sp_Task { sp_Context *context; Tile tile; } sp_Task;
#endif
/* (`struct WorkQueue`: a reference-side unit may include this header before work_queue.h) */
u32 sp_b200_AddRayTracingWorkQueue(struct WorkQueue *workQueue, sp_Context *ctx);
u32 sp_b200_DrainRayTracingWorkQueue(struct WorkQueue *queue, sp_Metrics *metricsBuffer, u32 maxMetrics);

/* =============================== additions (sp_b200_*) =============================== */

/* Batched forms of the two queries the reference's own performance tests time
 * (perf_tests/perf_tests.cpp:51-118 TestBvh: bvh_IntersectRay; :212-305 TestMeshMidphase:
 * sp_RayIntersectMesh), one ray per GPU thread, object-space rays against `mesh` (built by
 * sp_BuildMeshMidphase).  Leaves: countXorSum[3 q + 0] = leaves whose own box ray q passes (what
 * bvh_IntersectRay returns in `count`, without its 2048 cap), [1] = XOR of primitiveIndex * 0x9E3779B1,
 * [2] = sum of the primitive indices: a fingerprint of the leaf set.  Mesh: t (-1 on a miss) and the
 * triangle index.  kernelMs (optional): CUDA-event time of the kernel alone.  0 on success. */
int sp_b200_MeshIntersectedLeavesBatch(sp_Mesh mesh, u32 count, const vec3 *rayOrigins, const vec3 *rayDirections,
                                       u32 *countXorSum, f32 *kernelMs);
int sp_b200_RayIntersectMeshBatch(sp_Mesh mesh, u32 count, const vec3 *rayOrigins, const vec3 *rayDirections, f32 *t,
                                  i32 *triangleIndex, f32 *kernelMs);

/* simd_RayIntersectAabb4 (simd.h:198-271) as the DEVICE evaluates it, for known-answer tests: `count`
 * queries of four boxes (boxMin / boxMax: count x 4 x 3 floats) against (rayOrigin, invRayDirection)
 * taken exactly as the reference's function takes them.  masks[q * 3 + 0]: the exact form with the
 * SSE NaN semantics (second operand of min / max wins), [1]: the hardware min / max form used when no
 * product can be NaN, [2]: the conservative test of the resumable traversal (must contain [0]), or
 * 0xFFFFFFFF for a ray that traversal hands to the exact walk; bit k = box k.  tnear (optional):
 * count x 4 entry distances of the exact form. */
int sp_b200_RayIntersectAabb4Batch(u32 count, const f32 *boxMin, const f32 *boxMax, const vec3 *rayOrigins,
                                   const vec3 *invRayDirections, u32 *masks, f32 *tnear);


typedef void (*sp_b200_LogFn)(const char *message);

enum { SP_B200_ENV_NEAREST = 0, SP_B200_ENV_BILINEAR = 1 };
/* MOLLER_TRUMBORE: the reference's test, operation for operation (ray_intersection.cpp:156-190): the
 * parity default.  WATERTIGHT: Woop / Benthin / Wald 2013 with the reference's acceptance rules (front
 * faces, t > 0): no ray through a shared edge or vertex of two triangles can miss both.  It decides
 * those edge cases differently from the reference BY DESIGN, so images can differ from the
 * reference's in the pixels such rays belong to.  Every box test on the way is padded in this mode
 * (a ray that grazes a shared edge also grazes a face of both triangles' boxes).  Implemented in the
 * non-resumable walk: sp_b200_Render* use the per-pixel kernel while it is selected, whatever
 * renderMode says; ray queries and sp_PathTraceTile honour it directly. */
enum { SP_B200_TRIANGLE_MOLLER_TRUMBORE = 0, SP_B200_TRIANGLE_WATERTIGHT = 1 };
enum { SP_B200_MATH_F64_ROUNDED = 0, SP_B200_MATH_FAST_F32 = 1 };
/* WAVEFRONT: ray generation / traversal / shading as separate kernels over compact device
 * queues with warp-level lane refill (the default).  PER_PIXEL: one thread walks all samples and
 * bounces of a pixel (kept as a cross-check; both give bit-identical images). */
enum { SP_B200_RENDER_WAVEFRONT = 0, SP_B200_RENDER_PER_PIXEL = 1 };

/* Run-time form of the reference's compile-time knobs (config.h:15-33). */
typedef struct sp_b200_Params {
    u32 samplesPerPixel; /* SAMPLES_PER_PIXEL, config.h:23 (default 1024 there; 1 here) */
    u32 bounceCount;     /* literal 3 at simd_path_tracer.cpp:195; 1..8 */
    f32 radianceClamp;   /* RADIANCE_CLAMP, config.h:33; 0 disables */
    u32 envFilter;       /* SP_B200_ENV_*: the reference samples nearest (image.h:3-18) */
    u32 mathMode;        /* SP_B200_MATH_*: libm stand-ins, see DESIGN.md */
    u32 cullByDistance;  /* 1: ordered traversal with t-culling; 0: visit every intersected leaf
                            exactly like bvh_IntersectRay (bvh.cpp:203-311) */
    u32 tileWidth;       /* TILE_WIDTH / TILE_HEIGHT, config.h:15-16 (cost accounting granularity) */
    u32 tileHeight;
    u32 renderMode;      /* SP_B200_RENDER_*: how sp_b200_Render* schedules the work on the GPU */
    u32 samplesPerPass;  /* wavefront mode: samples per pixel traced per pass; 0 = automatic (all
                            of them, in bands of rows sized by sp_b200_SetPathsPerPass) */
    u32 triangleTest;    /* SP_B200_TRIANGLE_*: the ray / triangle test of every traversal */
} sp_b200_Params;

/* Kernel-side counters of the most recent launch (for the roofline). */
typedef struct sp_b200_Stats {
    u64 rays;
    u64 nodeVisits;     /* 4-wide internal nodes fetched (128 B each) */
    u64 triangleTests;  /* leaf triangles fetched and tested (48 B each) */
    u64 objectTests;    /* instances entered (ray transformed to object space) */
    u64 envClampedLookups; /* env texel indices past the image end (unclamped in image.h:3-18) */
    f32 kernelMs;       /* CUDA-event time of all kernels of the last call */
    f32 totalMs;        /* CUDA-event time of the whole call on the device (copies included) */
    f32 traceMs;        /* wavefront mode: summed CUDA-event time of the traversal kernel's launches
                           (the dominant kernel); otherwise = kernelMs */
    u32 traceLaunches;  /* how many launches traceMs covers */
    u64 tracedRays;     /* rays that went through the traversal kernel (rays minus sky-kernel samples) */
} sp_b200_Stats;

int sp_b200_Init(int device);            /* selects the device; 0 on success */
/* Several GPUs of the box in ONE process (SURVEY.md §8b "sp_b200_Init(deviceMask)", §8e): bit i of
 * the mask selects CUDA device i; the lowest one becomes the primary (every entry point keeps
 * running there), each further one gets its own worker thread, stream set and working buffers.
 * From then on sp_b200_RenderFrame / sp_b200_RenderFrameToDevice split the frame into strips of rows
 * cut at multiples of sp_b200_Params::tileHeight, one per device, all rendered concurrently; strip
 * boundaries are re-cut after every frame from the measured per-row cost (sp_b200_PartitionRows).
 * Scenes must be built (sp_BuildSceneBroadphase) AFTER this call: the further devices upload the
 * host copy it keeps, on first use.  Returns the number of devices in use, or -1. */
int sp_b200_InitDevices(u32 deviceMask);
/* The same with an explicit list (first entry = primary).  An ordinal may appear more than once:
 * every entry gets its own worker, streams and buffers, which exercises the whole multi-device path
 * -- strips, concurrent renders, gather, re-cut -- on a box with a single GPU. */
int sp_b200_InitDeviceList(const i32 *devices, u32 count);
u32 sp_b200_DeviceCount(void);
/* The frame gathered on the primary device: every other device sends its strip with a peer copy
 * (NVLink where the topology has it) into devicePixels, width x height RGBA f32 on the primary.
 * sp_b200_RenderFrame is the host-memory form: each device copies its own rows to the image plane's
 * (pinned) pixels over its own PCIe link. */
int sp_b200_RenderFrameToDevice(sp_Context *ctx, u32 frame, void *devicePixels, sp_Metrics *metrics);
/* Figures of device `index` (0 = primary) for the last multi-device frame and the rows it rendered;
 * 0 on success.  sp_b200_GetLastStats holds the frame's totals (times: the slowest device's). */
int sp_b200_GetDeviceStats(u32 index, sp_b200_Stats *stats, u32 *rowBegin, u32 *rowEnd);
/* The cut itself (host arithmetic; also what a multi-process host uses between its ranks): `parts`
 * contiguous strips of `height` rows with boundaries at multiples of `quantum`, minimising the
 * largest strip's summed cost (rowCost: one value per quantum row; NULL = even split).
 * bounds: parts + 1 ascending rows.  sp_b200_RowSeconds turns what the parts measured -- cost units
 * per quantum row (sp_b200_RenderRows tileRowCost) and seconds per part -- into seconds per row. */
void sp_b200_PartitionRows(u32 height, u32 quantum, u32 parts, const f64 *rowCost, u32 *bounds);
void sp_b200_RowSeconds(u32 height, u32 quantum, u32 parts, const u32 *bounds, const f64 *units,
                        const f64 *seconds, f64 *rowSeconds);
void sp_b200_Shutdown(void);             /* frees every device and host object */
void sp_b200_SetLogCallback(sp_b200_LogFn fn);
void sp_b200_SetStream(void *cudaStream);/* stream all later launches use (0 = default) */
void sp_b200_DefaultParams(sp_b200_Params *params);
void sp_b200_SetParams(const sp_b200_Params *params);
void sp_b200_GetParams(sp_b200_Params *params);
void sp_b200_GetLastStats(sp_b200_Stats *stats);
/* Collect node/triangle counters in the next launches (slower kernels; off by default). */
void sp_b200_EnableStats(int enable);
/* Kernels launched by this library since it was loaded (a count of real launches). */
u64 sp_b200_KernelLaunchCount(void);
/* Forget device copies of HdrImage pixel buffers (they are cached by host pointer). */
void sp_b200_FlushTextureCache(void);
/* Wavefront mode: paths kept in flight per pass (band height x samples per pass are derived from
 * it); 0 restores the default (32 Mi).  Results do not depend on it. */
void sp_b200_SetPathsPerPass(u32 paths);
/* Wavefront mode: pixels outside the padded screen rectangle of the scene's world bounds are
 * evaluated by a queue-less kernel (their rays cannot hit anything).  On by default; results do
 * not depend on it. */
void sp_b200_SetSkyCulling(int enable); /* 0 off; 1 on, per-sample loop for every sky pixel; 2 (default) on,
                                           sky pixels whose samples provably read one texel settled by one lookup */
/* Wavefront mode: the rays leaving the primary hits of a tile of 2048 paths are written in
 * direction order, so a warp of the trace kernel walks rays with neighbouring origins and similar
 * directions.  On by default; results do not depend on it. */
void sp_b200_SetRaySorting(int enable);
/* Wavefront mode, single-object scenes: the tree is walked once per pixel (padded centre ray, no
 * culling) to list the triangles any of the pixel's camera rays can meet; every sample then
 * evaluates only those, with the walk's own tests.  On by default; results do not depend on it. */
void sp_b200_SetPrimaryCandidates(int enable);
/* Copy-engine overlap (on by default; results do not depend on it).  HdrImage pixels (pinned host
 * memory) are uploaded on a second stream and the render stream waits for them only where the first
 * kernel that samples a texture starts -- behind the coverage pass, the candidate lists and the first
 * primary trace -- and sp_b200_RenderFrame / sp_b200_RenderRows copy finished rows to a pinned host
 * image band by band while later bands render.  0: one upload in front of the frame and one copy
 * behind it, both on the render stream (round 1's behaviour). */
void sp_b200_SetCopyOverlap(int enable);
/* The device already holds a copy of the HdrImage whose pixels == hostPixels (width x height RGBA
 * f32 at devicePixels, owned by the caller): use it instead of uploading.  readyEvent (a cudaEvent_t,
 * or NULL) is what the render stream waits for before the first kernel that samples a texture -- the
 * copy may still be in flight, e.g. an environment map of which every rank of a multi-process host
 * uploaded one slice and an NCCL all-gather over NVLink is assembling the rest.  devicePixels = NULL
 * forgets the association; sp_b200_FlushTextureCache forgets all of them. */
void sp_b200_SetDeviceTexture(const f32 *hostPixels, const void *devicePixels, u32 width, u32 height, void *readyEvent);
/* Wavefront mode tuning: a warp of the trace kernel retires and refills its lanes when fewer than
 * this many are still walking (1 = the whole warp starts and ends together), for primary rays,
 * direction-sorted bounce rays and all other rays; 0 keeps the default (1; measured choice of 1 or 12 per
 * scene; 12 for one object, 16 for several). */
void sp_b200_SetRefillThresholds(u32 primary, u32 sorted, u32 other);
/* Wavefront mode tuning: 1 = a ray that escapes is shaded by the trace kernel where it retires it (background
 * radiance, path folded back, radiance slot written) instead of going through the miss queue and a second kernel;
 * 0 = off; -1 (default) = the library's choice, currently off (measured, profiles/r2/README.md: C3 frame -1.8 %
 * with it on, C5 +0.8 %).  Same functions, same bits. */
void sp_b200_SetMissFusion(int mode);
/* Wavefront mode tuning: straggler eviction of the bounce traces.  A warp of the trace kernel walks a
 * packet of 32 rays; the rays of a packet end at different times, and the last few would keep the warp
 * busy at a fraction of its lanes.  With a threshold > 0, a packet whose walking lanes drop below it
 * parks them (a 128-byte continuation record each: traversal state and live stack entries) and the warp
 * starts the next packet; a second launch of the kernel refills its lanes from the parked records, so the
 * stragglers of many packets walk on together.  Every ray takes the same sequence of traversal steps
 * either way: results are bit-identical.  `sorted`: threshold for the direction-sorted launch (the first
 * bounce), `other`: for the later bounces; 0 = off (then sp_b200_SetRefillThresholds applies).  Until this is
 * called the thresholds follow the scene: 8 / off for one object, 16 / 16 for several (measured on C3 and C5). */
void sp_b200_SetStragglerEviction(u32 sorted, u32 other);
/* `count` draws of the reference's XorShift32 (math_utils.h:184-196) continuing *state (host side,
 * integer only): what seeded inputs like the reference's perf tests' are generated from. */
void sp_b200_XorShift32Stream(u32 *state, u32 count, u32 *values);
/* Progressive accumulation across frames (the reference's own to-do, main.cpp:75): deviceAccum
 * (pixelCount RGBA f32 on the device, owned by the caller) becomes the running mean of the frames
 * folded in so far, accum += (frame - accum) / (framesAccumulated + 1) per colour component in f32
 * (framesAccumulated = 0 copies the frame).  The new frame comes from deviceFrame or, if that is
 * NULL, from hostFrame; hostAccumOut (optional) receives the updated mean.  With sp_b200_RenderFrame's
 * per-(pixel, sample, frame) seeds, N frames of S samples accumulate to N x S independent samples. */
int sp_b200_AccumulateFrame(void *deviceAccum, const void *deviceFrame, const f32 *hostFrame, u32 pixelCount,
                            u32 framesAccumulated, f32 *hostAccumOut);
/* Seed of the per-(pixel, sample, frame) XorShift32 stream used by sp_b200_Render*. */
u32 sp_b200_Seed(u32 pixelIndex, u32 sample, u32 frame);

/* Whole-frame render with per-(pixel,sample) seeding: rows [rowBegin,rowEnd) of the image plane.
 * hostPixels (RGBA f32, full-image indexing, may be NULL) receives the rows by D2H copy;
 * devicePixels (device pointer to a full image, may be NULL) receives them in place.
 * tileRowCost (may be NULL): cost of every tile row (tileHeight rows; rowEnd-rowBegin rounded up) in
 * NANOSECONDS of this call's kernels: the work the kernels counted in the row -- wavefront mode: sky-
 * kernel samples, escaped rays (16 units) and surface hits (80 units) of the queue kernels; per-pixel
 * mode: rays -- converted class by class with the time that class of kernels took in this call, so the
 * rows add up to sp_b200_Stats::kernelMs.  What sp_b200_PartitionRows cuts by.  Returns 0 on success. */
int sp_b200_RenderRows(sp_Context *ctx, u32 rowBegin, u32 rowEnd, u32 frame, f32 *hostPixels,
                       void *devicePixels, sp_Metrics *metrics, u64 *tileRowCost);
/* The same call in two halves, so that a host which renders frame after frame (the reference's main
 * loop, main.cpp:1529-1557) keeps the device busy across the frame boundary.  Begin enqueues the whole
 * strip -- coverage pass, kernels, read-back of the counters, rows to the (pinned) host image band by
 * band -- and returns without waiting; End waits for that frame and fills metrics, sp_b200_GetLastStats
 * and tileRowCost (wantTileRowCost must have been set).  Up to TWO frames may be in flight: begin frame
 * k + 1, then end frame k -- the host-side epilogue of k and prologue of k + 1 then run under k + 1's
 * kernels, and the tail of k's row copies under them too.  Frames in flight must not share their
 * destination: with devicePixels = NULL each has its own device image; a caller that passes devicePixels
 * or hostPixels alternates between two of each.  The scene, camera and material system of ctx must stay
 * as they are until End (a scene may be REBUILT with identical content in between: the device copy a frame
 * in flight uses is kept until it ends).  Begin returns the frame's slot (0 or 1) for End, -1 if two
 * frames are already in flight.  sp_b200_RenderRows = Begin + End. */
int sp_b200_RenderRowsBegin(sp_Context *ctx, u32 rowBegin, u32 rowEnd, u32 frame, f32 *hostPixels,
                            void *devicePixels, int wantTileRowCost);
int sp_b200_RenderRowsEnd(int slot, sp_Metrics *metrics, u64 *tileRowCost);
/* All rows into ctx->camera->imagePlane->pixels (host). */
int sp_b200_RenderFrame(sp_Context *ctx, u32 frame, sp_Metrics *metrics);
/* Asset input (the step before the path; replaces LoadMesh, src/mesh.cpp:5-62, whose assimp
 * import is not available): Wavefront OBJ -> the VertexPNT[] / u32[] pair sp_CreateMesh takes.
 * One vertex per distinct (v, vt, vn) corner in first-seen order, polygons fan-triangulated in
 * file order (triangle i of the result is the i-th triangle of the file).  Same fields as the
 * reference's MeshData (src/mesh.h:24-30).  Returns 1 on success, 0 on a missing / malformed
 * file (house convention of LoadExrImage, src/asset_loader/asset_loader.h:11-22); the arrays are
 * malloc'ed: release with sp_b200_FreeMeshData. */
typedef struct sp_b200_MeshData {
    VertexPNT *vertices;
    u32 *indices;
    u32 vertexCount;
    u32 indexCount;
} sp_b200_MeshData;
int sp_b200_LoadObj(const char *path, sp_b200_MeshData *out);
void sp_b200_FreeMeshData(sp_b200_MeshData *mesh);

/* OpenEXR input behind the reference's own loader ABI (src/asset_loader/asset_loader.h:11-22; there
 * implemented with the vendored tinyexr): 0 on success, 1 on failure; pixels = malloc'ed RGBA f32,
 * rows top to bottom, alpha 1 when the file has none, a single channel replicated; the caller
 * free()s them.  Single-part scanline files, HALF / FLOAT channels, compression NONE / RLE / ZIPS /
 * ZIP / PIZ, scanline or single-part tiled (level 0); anything else (multi-part, deep, PXR24, B44,
 * DWA -- which the reference's tinyexr refuses too) is refused with 1. */
int LoadExrImage(HdrImage *image, const char *path);

/* Output stage (the step after the path, src/shaders/post_processing.frag.glsl:19-26
 * PerformToneMapping): color *= exposure; color = color / (1 + color); pow(color, 1/2.2); then the
 * 8-bit UNORM store of the colour attachment, alpha 255, r in the low byte (the layout of ToColor,
 * src/math_lib.h:523-532).  pixelCount RGBA f32 pixels from devicePixels if non-NULL, else from
 * hostPixels; RGBA8 result to hostRGBA8 and/or deviceRGBA8 (either may be NULL, not both). */
int sp_b200_ToneMap(const f32 *hostPixels, const void *devicePixels, u32 pixelCount, f32 exposure,
                    u32 *hostRGBA8, void *deviceRGBA8);

/* Image files out of the path (the reference has no writer; its frames go to the swap chain,
 * main.cpp:1607).  House convention of LoadExrImage: 0 on success, 1 on failure.
 * sp_b200_SaveExrImage: RGBA f32 image -> single-part scanline OpenEXR file, channels A B G R,
 *   pixelType HALF (round to nearest even) or FLOAT, compression NONE / ZIPS / ZIP; what LoadExrImage
 *   (this library's and the reference's tinyexr-based one) reads back bit for bit.
 * sp_b200_SavePpm: RGBA8 pixels as sp_b200_ToneMap stores them (r in the low byte) -> binary PPM. */
#define SP_B200_EXR_HALF 1u
#define SP_B200_EXR_FLOAT 2u
#define SP_B200_EXR_NONE 0u
#define SP_B200_EXR_ZIPS 2u
#define SP_B200_EXR_ZIP 3u
int sp_b200_SaveExrImage(const HdrImage *image, const char *path, u32 pixelType, u32 compression);
/* the same as a single-part tiled file (one level, tileWidth x tileHeight tiles, a chunk per tile) */
int sp_b200_SaveExrImageTiled(const HdrImage *image, const char *path, u32 pixelType, u32 compression,
                              u32 tileWidth, u32 tileHeight);
int sp_b200_SavePpm(const u32 *rgba8, u32 width, u32 height, const char *path);

/* Environment pre-processing on the GPU (src/cubemap.cpp; the reference bakes both maps on the CPU
 * at start-up, main.cpp:1307-1315).  Faces are written layer-major in the reference's order
 * (+X -X +Y -Y +Z -Z, cubemap.cpp:13-21, basis vectors :54-104), each width x height RGBA f32, rows
 * top to bottom: 6 * width * height * 4 floats, to hostFaces and/or deviceFaces (either may be
 * NULL, not both).  The source goes through the same device texture cache as registered textures.
 *
 * sp_b200_CreateCubeMap replaces CreateCubeMap(HdrImage, MemoryArena*, u32, u32)
 * (cubemap.cpp:237-291): per texel Normalize(forward + right*fx + up*fy) -> ToSphericalCoordinates
 * -> MapToEquirectangular -> v flipped -> SampleImageBilinear (image.h:34-73), all four channels.
 *
 * sp_b200_CreateIrradianceCubeMap replaces CreateIrradianceCubeMap(HdrImage, MemoryArena*, u32,
 * u32, u32 samplesPerPixel = 32) (cubemap.cpp:108-233).  The reference picks its sampling at
 * compile time (IRRADIANCE_CUBEMAP_USE_UNIFORM_SAMPLING, config.h:47, = 1); here it is an argument:
 *   SP_B200_IRRADIANCE_UNIFORM  the (phi, theta) grid of cubemap.cpp:152-198; sampleDelta is the
 *       loop step (0.1f in the reference, "TODO: Parameterize"); samplesPerPixel is ignored, as there;
 *   SP_B200_IRRADIANCE_RANDOM   cubemap.cpp:200-224: samplesPerPixel jittered directions per texel
 *       drawn from ONE serial XorShift32 stream over all texels (state 0x45BA12F3, :122-123); every
 *       texel's place in that stream is reached by a GF(2) jump, so the result equals the serial loop's.
 * Each texel's terms are summed in the reference's order; radiance clamp = RADIANCE_CLAMP (config.h:36).
 * Both return 0 (asserts abort through LogMessage like the rest of the library). */
#define SP_B200_IRRADIANCE_UNIFORM 0u
#define SP_B200_IRRADIANCE_RANDOM 1u
int sp_b200_CreateCubeMap(const HdrImage *equirect, u32 width, u32 height, f32 *hostFaces,
                          void *deviceFaces);
int sp_b200_CreateIrradianceCubeMap(const HdrImage *equirect, u32 width, u32 height, u32 samplesPerPixel,
                                    u32 sampling, f32 sampleDelta, f32 *hostFaces, void *deviceFaces);
/* Primary-ray closest hits of (sample, frame): per pixel triangle index (-1 = miss), object
 * index and world t.  Output arrays are host memory, any may be NULL. */
int sp_b200_PrimaryHits(sp_Context *ctx, u32 sample, u32 frame, i32 *triangleIndex,
                        i32 *objectIndex, f32 *t);
/* Batched sp_RayIntersectScene: origins/directions are n x vec3 host arrays. */
int sp_b200_RayIntersectSceneBatch(sp_Scene *scene, u32 count, const vec3 *rayOrigins,
                                   const vec3 *rayDirections, sp_RayIntersectSceneResult *results,
                                   i32 *triangleIndex, i32 *objectIndex, sp_Metrics *metrics);
/* bvh_IntersectRay semantics on this library's tree (bvh.cpp:203-311): collects the indices of
 * every leaf whose box chain the ray passes.  Returns the count; *errorOccurred is set when
 * more than maxIntersections leaves are hit. */
u32 sp_b200_MeshIntersectedLeaves(sp_Mesh mesh, vec3 rayOrigin, vec3 rayDirection,
                                  u32 *leafIndices, u32 maxIntersections, b32 *errorOccurred);
/* Which builder sp_BuildMeshMidphase uses (replaces bvh_CreateTree, bvh.cpp:51-200).  HOST_SAH
 * (default): binned / swept SAH on the host.  DEVICE_LBVH: Morton keys, radix sort, binary radix
 * tree and bottom-up boxes on the GPU (Karras 2012), then the same 4-wide collapse on the host; it
 * falls back to HOST_SAH for meshes of fewer than 8 triangles, for trees deeper than the traversal
 * stack and on CUDA errors.  Results of ray queries and renders do not depend on the builder (every
 * triangle sits alone in a child slot with its own AABB either way); traversal cost does. */
#define SP_B200_BUILDER_HOST_SAH 0u
#define SP_B200_BUILDER_DEVICE_LBVH 1u
typedef struct sp_b200_BuildInfo {
    u32 builder;       /* builder selected for the last sp_BuildMeshMidphase */
    u32 fellBack;      /* 1 if DEVICE_LBVH was selected but the host builder made the tree */
    u32 triangleCount; u32 nodeCount; u32 maxDepth; u32 stackNeed;
    f32 deviceMs;      /* CUDA-event time of the four device passes */
    f32 wallMs;        /* whole call: snapshot, boxes, build, collapse */
} sp_b200_BuildInfo;
void sp_b200_SetMeshBuilder(u32 builder);
void sp_b200_GetLastBuildInfo(sp_b200_BuildInfo *info);
/* Host-side structure queries on the midphase tree (test hooks, cf. unit_tests/test_bvh.cpp). */
typedef struct sp_b200_TreeInfo {
    u32 leafCount; u32 nodeCount; u32 maxDepth; b32 allLeavesReachable;
    b32 parentsContainChildren; vec3 rootMin; vec3 rootMax;
} sp_b200_TreeInfo;
void sp_b200_MeshTreeInfo(sp_Mesh mesh, sp_b200_TreeInfo *info);
/* Bytes sp_BuildSceneBroadphase uploaded for this scene (nodes, triangles, shading data, instances). */
u64 sp_b200_SceneDeviceBytes(sp_Scene *scene);
void sp_b200_ReleaseMesh(sp_Mesh *mesh);
void sp_b200_ReleaseScene(sp_Scene *scene);

#ifdef __cplusplus
}
#endif

#ifndef SP_B200_USE_REFERENCE_TYPES
#if defined(__cplusplus)
#define SP_B200_SIZE_CHECK(T, N) static_assert(sizeof(T) == (N), #T " layout differs from the reference")
#else
#define SP_B200_SIZE_CHECK(T, N) _Static_assert(sizeof(T) == (N), #T " layout differs from the reference")
#endif
SP_B200_SIZE_CHECK(vec3, 12);
SP_B200_SIZE_CHECK(vec4, 16);
SP_B200_SIZE_CHECK(mat4, 64);
SP_B200_SIZE_CHECK(VertexPNT, 32);
SP_B200_SIZE_CHECK(sp_Mesh, 64);
SP_B200_SIZE_CHECK(sp_Scene, 7104);
SP_B200_SIZE_CHECK(sp_Material, 36);
SP_B200_SIZE_CHECK(sp_PathVertex, 60);
SP_B200_SIZE_CHECK(sp_MaterialSystem, 1616);
SP_B200_SIZE_CHECK(HdrImage, 16);
SP_B200_SIZE_CHECK(sp_Camera, 88);
SP_B200_SIZE_CHECK(Tile, 16);
SP_B200_SIZE_CHECK(sp_Metrics, 96);
SP_B200_SIZE_CHECK(sp_RayIntersectSceneResult, 28);
SP_B200_SIZE_CHECK(RayIntersectTriangleResult, 24);
#undef SP_B200_SIZE_CHECK
#endif

#endif /* SP_B200_H */
