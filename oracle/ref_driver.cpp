// oracle/ref_driver.cpp -- harness around the UNMODIFIED reference sources.
//
// TEST INFRASTRUCTURE ONLY (see oracle/ora_api.h).  This translation unit #includes the
// reference's own .h/.cpp files from /root/reference/src in the reference's unity-build style
// (pattern: perf_tests/perf_tests.cpp:5-31, unit_tests/test_simd_path_tracer.cpp:1-31) and adds
// a C harness around them.  No reference source is copied into this repository; the output is
// oracle/_ref/libspref.so (git-ignored, built by oracle/Makefile).
//
// The only knob applied from outside is SAMPLES_PER_PIXEL, which the reference reads through a
// macro (simd_path_tracer.cpp:194); the reference's own tests redefine it the same way
// (test_simd_path_tracer.cpp:24).  Bounce count is a literal 3 in the reference (:195).
//
// Build flags: -O2, no -march=native / -mfma / -ffast-math (SURVEY.md §7 hard part 1).

#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstring>
#include <math.h>
#include <thread>
#include <vector>

// Second oracle mode, "deterministic math" (oracle/_ref/libspref_dm.so): the same unmodified
// reference sources, but the four single-precision libm calls on the path (sinf, cosf, atan2f,
// powf -- reached through math_utils.h:63-97) are redirected to "evaluate in double, round once
// to float".  glibc's float functions are not correctly rounded (they differ from the rounded
// double result for 1.3 % / 8 % / 0.07 % of inputs, DESIGN.md "libm"), and a GPU cannot call
// glibc; with the redirect both sides compute the same function, so everything else on the path
// (traversal, shading, accumulation) can be compared bit-for-bit.  The primary mode
// (libspref.so) keeps glibc's functions and is compared within a stated tolerance.
#ifdef ORA_DETERMINISTIC_MATH
static inline float ora_dm_sinf(float x) { return (float)sin((double)x); }
static inline float ora_dm_cosf(float x) { return (float)cos((double)x); }
static inline float ora_dm_atan2f(float y, float x) { return (float)atan2((double)y, (double)x); }
static inline float ora_dm_powf(float x, float y)
{
    if (y == 5.0f)
    {
        double d = (double)x;
        double d2 = d * d;
        return (float)(d2 * d2 * d);
    }
    return (float)pow((double)x, (double)y);
}
#define sinf ora_dm_sinf
#define cosf ora_dm_cosf
#define atan2f ora_dm_atan2f
#define powf ora_dm_powf
#define ORA_NAME "reference-dm"
#else
#define ORA_NAME "reference"
#endif

// --- the reference, verbatim -------------------------------------------------------------
#include "config.h"
#include "platform.h"
#include "math_lib.h"
#include "tile.h"
#include "memory_pool.h"
#include "bvh.h"
#include "ray_intersection.h"
#include "asset_loader/asset_loader.h"
#include "image.h"
#include "mesh.h"
#include "sp_scene.h"
#include "sp_material_system.h"
#include "simd_path_tracer.h"
#include "sp_metrics.h"
#include "simd.h"
#include "aabb.h"
#include "intrinsics.h"
#include "work_queue.h"

#include "memory_pool.cpp"
#include "ray_intersection.cpp"

#undef SAMPLES_PER_PIXEL
static u32 g_oraSamplesPerPixel = 1;
#define SAMPLES_PER_PIXEL g_oraSamplesPerPixel

#include "bvh.cpp"
#include "sp_scene.cpp"
#include "sp_material_system.cpp"
#include "simd_path_tracer.cpp"
// src/cubemap.cpp twice: once as config.h:47 configures it (uniform irradiance sampling) and once
// with the switch flipped, so the random branch (cubemap.cpp:200-224) is the reference's code too.
// The file defines only `internal` functions and its own small types; a namespace per copy.
namespace ora_cubemap_uniform {
#include "cubemap.cpp"
}
#undef IRRADIANCE_CUBEMAP_USE_UNIFORM_SAMPLING
#define IRRADIANCE_CUBEMAP_USE_UNIFORM_SAMPLING 0
namespace ora_cubemap_random {
#include "cubemap.cpp"
}
#undef IRRADIANCE_CUBEMAP_USE_UNIFORM_SAMPLING
#define IRRADIANCE_CUBEMAP_USE_UNIFORM_SAMPLING 1
// -----------------------------------------------------------------------------------------

#include "ora_api.h"

static void OraLog(const char *fmt, ...)
{
    va_list args;
    va_start(args, fmt);
    vfprintf(stderr, fmt, args);
    fputc('\n', stderr);
    va_end(args);
}

struct OraInit
{
    OraInit() { LogMessage = &OraLog; }
};
static OraInit g_oraInit;

struct ora_Scene
{
    sp_Scene scene;
    sp_MaterialSystem materials;
    sp_Camera camera;
    ImagePlane imagePlane;
    sp_Context ctx;
    std::vector<sp_Mesh> meshes;
    std::vector<void *> allocations;
    MemoryArena sceneArena;
};

static void *OraAlloc(ora_Scene *s, size_t bytes)
{
    void *p = calloc(1, bytes ? bytes : 1);
    s->allocations.push_back(p);
    return p;
}

extern "C" const char *ora_name(void) { return ORA_NAME; }
extern "C" uint32_t ora_max_bounces(void) { return 3; }

extern "C" ora_Scene *ora_create(void)
{
    ora_Scene *s = new ora_Scene();
    memset(&s->scene, 0, sizeof(s->scene));
    memset(&s->materials, 0, sizeof(s->materials));
    memset(&s->camera, 0, sizeof(s->camera));
    memset(&s->imagePlane, 0, sizeof(s->imagePlane));
    size_t arenaBytes = Megabytes(2);
    InitializeMemoryArena(&s->sceneArena, OraAlloc(s, arenaBytes), arenaBytes);
    sp_InitializeScene(&s->scene, &s->sceneArena);
    s->camera.imagePlane = &s->imagePlane;
    s->ctx.camera = &s->camera;
    s->ctx.scene = &s->scene;
    s->ctx.materialSystem = &s->materials;
    return s;
}

extern "C" void ora_destroy(ora_Scene *s)
{
    if (!s) return;
    for (void *p : s->allocations) free(p);
    delete s;
}

extern "C" int ora_add_mesh(ora_Scene *s, const float *vertices, uint32_t vertexCount,
                            const uint32_t *indices, uint32_t indexCount, uint32_t smooth)
{
    VertexPNT *v = (VertexPNT *)OraAlloc(s, sizeof(VertexPNT) * (size_t)vertexCount);
    u32 *idx = (u32 *)OraAlloc(s, sizeof(u32) * (size_t)indexCount);
    memcpy(v, vertices, sizeof(VertexPNT) * (size_t)vertexCount);
    memcpy(idx, indices, sizeof(u32) * (size_t)indexCount);

    sp_Mesh mesh = sp_CreateMesh(v, vertexCount, idx, indexCount, smooth);

    // bvh_CreateTree reserves count*10 nodes plus two pair buffers (bvh.cpp:57-68)
    size_t triCount = indexCount / 3;
    size_t accelBytes = triCount * 10 * sizeof(bvh_Node) +
                        2 * triCount * sizeof(bvh_NodeDistSqPair) + 4096;
    size_t tempBytes = 2 * triCount * sizeof(vec3) + 4096;
    MemoryArena accel, temp;
    InitializeMemoryArena(&accel, OraAlloc(s, accelBytes), accelBytes);
    InitializeMemoryArena(&temp, OraAlloc(s, tempBytes), tempBytes);
    sp_BuildMeshMidphase(&mesh, &accel, &temp);

    s->meshes.push_back(mesh);
    return (int)s->meshes.size() - 1;
}

extern "C" int ora_add_object(ora_Scene *s, uint32_t mesh, uint32_t material,
                              const float *p, const float *q, const float *sc)
{
    if (s->scene.objectCount >= SP_SCENE_MAX_OBJECTS) return -1;
    quat rotation = {q[0], q[1], q[2], q[3]};
    sp_AddObjectToScene(&s->scene, s->meshes[mesh], material, Vec3(p[0], p[1], p[2]), rotation,
                        Vec3(sc[0], sc[1], sc[2]));
    return (int)s->scene.objectCount - 1;
}

extern "C" void ora_build(ora_Scene *s)
{
    ResetMemoryArena(&s->scene.memoryArena);
    memset(&s->scene.broadphaseTree, 0, sizeof(s->scene.broadphaseTree));
    if (s->scene.objectCount > 0)
    {
        sp_BuildSceneBroadphase(&s->scene);
    }
}

extern "C" int ora_register_material(ora_Scene *s, uint32_t id, const float *albedo,
                                     uint32_t albedoTexture, const float *emission,
                                     uint32_t emissionTexture, float roughness)
{
    sp_Material m = {};
    m.albedo = Vec3(albedo[0], albedo[1], albedo[2]);
    m.albedoTexture = albedoTexture;
    m.emission = Vec3(emission[0], emission[1], emission[2]);
    m.emissionTexture = emissionTexture;
    m.roughness = roughness;
    return (int)sp_RegisterMaterial(&s->materials, m, id);
}

extern "C" int ora_register_texture(ora_Scene *s, uint32_t id, const float *pixels,
                                    uint32_t width, uint32_t height)
{
    // SampleImageNearest (image.h:3-18) does not clamp: uv.y == 1 (a ray leaving straight down,
    // which the theta-uniform hemisphere sampler produces about once per 10^4 bounce rays)
    // indexes row `height`, i.e. up to width+1 texels past the end of the caller's buffer.  In
    // the reference that is an out-of-bounds heap read.  The harness therefore hands the
    // reference a private copy of the image followed by width+1 texels equal to the last texel,
    // which makes the unmodified code return what the GPU's clamped lookup returns
    // (spb_core.cuh sample_nearest) instead of faulting.
    size_t texels = (size_t)width * height;
    size_t padded = texels + width + 1;
    float *copy = (float *)OraAlloc(s, padded * 4 * sizeof(float));
    memcpy(copy, pixels, texels * 4 * sizeof(float));
    for (size_t i = texels; i < padded && texels > 0; ++i)
    {
        memcpy(copy + i * 4, pixels + (texels - 1) * 4, 4 * sizeof(float));
    }
    HdrImage image = {};
    image.pixels = copy;
    image.width = width;
    image.height = height;
    return (int)sp_RegisterTexture(&s->materials, image, id);
}

extern "C" void ora_set_background(ora_Scene *s, uint32_t materialId)
{
    s->materials.backgroundMaterialId = materialId;
}

extern "C" void ora_configure_camera(ora_Scene *s, const float *p, const float *q,
                                     float filmDistance, uint32_t width, uint32_t height)
{
    s->imagePlane.width = width;
    s->imagePlane.height = height;
    s->imagePlane.pixels = NULL;
    quat rotation = {q[0], q[1], q[2], q[3]};
    sp_ConfigureCamera(&s->camera, &s->imagePlane, Vec3(p[0], p[1], p[2]), rotation,
                       filmDistance);
}

extern "C" uint32_t ora_seed(uint32_t pixelIndex, uint32_t sample, uint32_t frame)
{
    uint32_t h = pixelIndex * 0x9E3779B1u;
    h ^= sample * 0x85EBCA77u;
    h ^= frame * 0xC2B2AE3Du;
    h ^= h >> 16;
    h *= 0x7FEB352Du;
    h ^= h >> 15;
    h *= 0x846CA68Bu;
    h ^= h >> 16;
    return h | 1u;
}

static void OraAddMetrics(uint64_t *dst, const sp_Metrics &m)
{
    if (!dst) return;
    for (u32 i = 0; i < SP_MAX_METRICS; ++i) dst[i] += m.values[i];
}

static void OraRequireBounces(uint32_t bounces)
{
    if (bounces != 3)
    {
        OraLog("oracle/_ref: the reference hard-codes 3 bounces (simd_path_tracer.cpp:195); "
               "requested %u", bounces);
        abort();
    }
}

extern "C" void ora_render_seeded(ora_Scene *s, float *rgba, uint32_t x0, uint32_t y0,
                                  uint32_t x1, uint32_t y1, uint32_t spp, uint32_t bounces,
                                  uint32_t frame, uint32_t threads, uint64_t *metrics)
{
    OraRequireBounces(bounces);
    g_oraSamplesPerPixel = 1;
    u32 width = s->imagePlane.width;
    s->imagePlane.pixels = (vec4 *)rgba;
    if (threads == 0) threads = 1;

    std::vector<sp_Metrics> perThread(threads);
    memset(perThread.data(), 0, sizeof(sp_Metrics) * threads);

    auto worker = [&](u32 tid) {
        sp_Metrics *m = &perThread[tid];
        f32 weight = 1.0f / (f32)spp;
        for (u32 y = y0 + tid; y < y1; y += threads)
        {
            for (u32 x = x0; x < x1; ++x)
            {
                vec3 total = {};
                for (u32 sample = 0; sample < spp; ++sample)
                {
                    RandomNumberGenerator rng;
                    rng.state = ora_seed(x + y * width, sample, frame);
                    Tile tile = {x, y, x + 1, y + 1};
                    u64 cycles = m->values[sp_Metric_CyclesElapsed];
                    sp_PathTraceTile(&s->ctx, tile, &rng, m);
                    m->values[sp_Metric_CyclesElapsed] += cycles;
                    // With SAMPLES_PER_PIXEL == 1 the pixel holds radiance * (1/1)
                    vec3 radiance = s->imagePlane.pixels[x + y * width].xyz;
                    total += radiance * weight;
                }
                s->imagePlane.pixels[x + y * width] = Vec4(total, 1);
            }
        }
    };

    std::vector<std::thread> pool;
    for (u32 t = 1; t < threads; ++t) pool.emplace_back(worker, t);
    worker(0);
    for (auto &t : pool) t.join();
    for (u32 t = 0; t < threads; ++t) OraAddMetrics(metrics, perThread[t]);
    s->imagePlane.pixels = NULL;
}

// Restated from main.cpp:246-250,731-759,819-844 (main.cpp itself needs GLFW/Vulkan).
struct OraTask
{
    sp_Context *context;
    Tile tile;
};

extern "C" double ora_render_tiles(ora_Scene *s, float *rgba, uint32_t tileW, uint32_t tileH,
                                   uint32_t spp, uint32_t bounces, uint32_t threads,
                                   uint64_t *metrics)
{
    OraRequireBounces(bounces);
    g_oraSamplesPerPixel = spp;
    s->imagePlane.pixels = (vec4 *)rgba;
    if (threads == 0) threads = 1;

    u32 tilesX = (s->imagePlane.width + tileW - 1) / tileW;
    u32 tilesY = (s->imagePlane.height + tileH - 1) / tileH;
    u32 maxTiles = tilesX * tilesY;
    std::vector<Tile> tiles(maxTiles);
    u32 tileCount = ComputeTiles(s->imagePlane.width, s->imagePlane.height, tileW, tileH,
                                 tiles.data(), maxTiles);

    MemoryArena queueArena;
    size_t queueBytes = sizeof(OraTask) * (size_t)tileCount + 64;
    std::vector<u8> queueStorage(queueBytes);
    InitializeMemoryArena(&queueArena, queueStorage.data(), queueBytes);
    WorkQueue queue = CreateWorkQueue(&queueArena, sizeof(OraTask), tileCount);

    std::vector<sp_Metrics> tileMetrics(tileCount);
    memset(tileMetrics.data(), 0, sizeof(sp_Metrics) * tileCount);

    auto begin = std::chrono::steady_clock::now();
    for (u32 i = 0; i < tileCount; ++i)
    {
        OraTask task = {&s->ctx, tiles[i]};
        WorkQueuePush(&queue, &task, sizeof(task));
    }

    auto worker = [&]() {
        for (;;)
        {
            // Unlike the reference's forever-spinning workers, exit once the queue drains
            i32 index = AtomicExchangeAdd(&queue.head, 1);
            if (index >= queue.tail) break;
            OraTask *task = (OraTask *)((u8 *)queue.buffer + (size_t)index * sizeof(OraTask));
            RandomNumberGenerator rng = {};
            rng.state = 0xF51C0E49;
            sp_Metrics m = {};
            sp_PathTraceTile(task->context, task->tile, &rng, &m);
            tileMetrics[index] = m;
        }
    };
    std::vector<std::thread> pool;
    for (u32 t = 1; t < threads; ++t) pool.emplace_back(worker);
    worker();
    for (auto &t : pool) t.join();
    auto end = std::chrono::steady_clock::now();

    for (u32 i = 0; i < tileCount; ++i) OraAddMetrics(metrics, tileMetrics[i]);
    s->imagePlane.pixels = NULL;
    g_oraSamplesPerPixel = 1;
    return std::chrono::duration<double>(end - begin).count();
}

extern "C" double ora_render_tile_list(ora_Scene *s, float *rgba, const uint32_t *tileList,
                                       uint32_t count, uint32_t spp, uint32_t bounces,
                                       uint32_t threads, uint64_t *metrics)
{
    OraRequireBounces(bounces);
    g_oraSamplesPerPixel = spp;
    s->imagePlane.pixels = (vec4 *)rgba;
    if (threads == 0) threads = 1;

    MemoryArena queueArena;
    size_t queueBytes = sizeof(OraTask) * (size_t)count + 64;
    std::vector<u8> queueStorage(queueBytes);
    InitializeMemoryArena(&queueArena, queueStorage.data(), queueBytes);
    WorkQueue queue = CreateWorkQueue(&queueArena, sizeof(OraTask), count ? count : 1);
    std::vector<sp_Metrics> tileMetrics(count ? count : 1);
    memset(tileMetrics.data(), 0, sizeof(sp_Metrics) * tileMetrics.size());

    auto begin = std::chrono::steady_clock::now();
    for (u32 i = 0; i < count; ++i)
    {
        Tile tile = {tileList[i * 4 + 0], tileList[i * 4 + 1], tileList[i * 4 + 2], tileList[i * 4 + 3]};
        OraTask task = {&s->ctx, tile};
        WorkQueuePush(&queue, &task, sizeof(task));
    }
    auto worker = [&]() {
        for (;;)
        {
            i32 index = AtomicExchangeAdd(&queue.head, 1);
            if (index >= queue.tail) break;
            OraTask *task = (OraTask *)((u8 *)queue.buffer + (size_t)index * sizeof(OraTask));
            RandomNumberGenerator rng = {};
            rng.state = 0xF51C0E49; // main.cpp:738-739
            sp_Metrics m = {};
            sp_PathTraceTile(task->context, task->tile, &rng, &m);
            tileMetrics[index] = m;
        }
    };
    std::vector<std::thread> pool;
    for (u32 t = 1; t < threads; ++t) pool.emplace_back(worker);
    worker();
    for (auto &t : pool) t.join();
    auto end = std::chrono::steady_clock::now();
    for (u32 i = 0; i < count; ++i) OraAddMetrics(metrics, tileMetrics[i]);
    s->imagePlane.pixels = NULL;
    g_oraSamplesPerPixel = 1;
    return std::chrono::duration<double>(end - begin).count();
}

extern "C" void ora_path_trace_tile(ora_Scene *s, float *rgba, uint32_t minX, uint32_t minY,
                                    uint32_t maxX, uint32_t maxY, uint32_t spp,
                                    uint32_t bounces, uint32_t *rngState, uint64_t *metrics)
{
    OraRequireBounces(bounces);
    g_oraSamplesPerPixel = spp;
    s->imagePlane.pixels = (vec4 *)rgba;
    RandomNumberGenerator rng = {*rngState};
    sp_Metrics m = {};
    Tile tile = {minX, minY, maxX, maxY};
    sp_PathTraceTile(&s->ctx, tile, &rng, &m);
    *rngState = rng.state;
    OraAddMetrics(metrics, m);
    s->imagePlane.pixels = NULL;
    g_oraSamplesPerPixel = 1;
}

// Winner bookkeeping the reference drops (sp_scene.cpp:164,189-196): the same calls in the same
// order as sp_RayIntersectScene / sp_RayIntersectMesh, additionally remembering which leaf won.
static f32 OraClosestHit(sp_Scene *scene, vec3 rayOrigin, vec3 rayDirection, i32 *outObject,
                         i32 *outTriangle)
{
    f32 bestT = -1.0f;
    *outObject = -1;
    *outTriangle = -1;

    bvh_Node *objectNodes[32] = {};
    bvh_IntersectRayResult broadphase = bvh_IntersectRay(&scene->broadphaseTree, rayOrigin,
        rayDirection, objectNodes, ArrayCount(objectNodes));
    Assert(!broadphase.errorOccurred);

    for (u32 i = 0; i < broadphase.count; ++i)
    {
        u32 objectIndex = objectNodes[i]->leafIndex;
        mat4 invModelMatrix = scene->invModelMatrices[objectIndex];
        mat4 modelMatrix = scene->modelMatrices[objectIndex];
        sp_Mesh mesh = scene->meshes[objectIndex];

        vec3 localRayOrigin = TransformPoint(rayOrigin, invModelMatrix);
        vec3 localRayDirection = Normalize(TransformVector(rayDirection, invModelMatrix));

        bvh_Node *leaves[128] = {};
        bvh_IntersectRayResult midphase = bvh_IntersectRay(&mesh.midphaseTree, localRayOrigin,
            localRayDirection, leaves, ArrayCount(leaves));
        Assert(!midphase.errorOccurred);

        f32 localT = -1.0f;
        i32 localTriangle = -1;
        for (u32 j = 0; j < midphase.count; ++j)
        {
            u32 triangleIndex = leaves[j]->leafIndex;
            vec3 a = mesh.vertices[mesh.indices[triangleIndex * 3 + 0]].position;
            vec3 b = mesh.vertices[mesh.indices[triangleIndex * 3 + 1]].position;
            vec3 c = mesh.vertices[mesh.indices[triangleIndex * 3 + 2]].position;
            RayIntersectTriangleResult hit =
                RayIntersectTriangle(localRayOrigin, localRayDirection, a, b, c);
            if (hit.t > 0.0f)
            {
                if (hit.t < localT || localT < 0.0f)
                {
                    localT = hit.t;
                    localTriangle = (i32)triangleIndex;
                }
            }
        }

        if (localT >= 0.0f)
        {
            vec3 localHitPoint = localRayOrigin + localRayDirection * localT;
            vec3 worldHitPoint = TransformPoint(localHitPoint, modelMatrix);
            f32 t = Dot(worldHitPoint - rayOrigin, rayDirection);
            if (t < bestT || bestT < 0.0f)
            {
                bestT = t;
                *outObject = (i32)objectIndex;
                *outTriangle = localTriangle;
            }
        }
    }
    return bestT;
}

extern "C" void ora_primary_hits(ora_Scene *s, int32_t *triId, int32_t *objId, float *tOut,
                                 float *rayDir3, uint32_t sample, uint32_t frame,
                                 uint32_t threads)
{
    u32 width = s->imagePlane.width;
    u32 height = s->imagePlane.height;
    sp_Camera *camera = &s->camera;
    if (threads == 0) threads = 1;

    auto worker = [&](u32 tid) {
        for (u32 y = tid; y < height; y += threads)
        {
            for (u32 x = 0; x < width; ++x)
            {
                RandomNumberGenerator rngStorage;
                RandomNumberGenerator *rng = &rngStorage;
                rng->state = ora_seed(x + y * width, sample, frame);

                // Same expressions as simd_path_tracer.cpp:219-230, same compiler, hence the
                // same (unspecified) evaluation order of the two RandomBilateral calls.
                vec2 pixelPosition = Vec2((f32)x, (f32)y) + Vec2(0.5);
                pixelPosition +=
                    Vec2(camera->halfPixelWidth * RandomBilateral(rng),
                        camera->halfPixelHeight * RandomBilateral(rng));
                vec3 filmP = {};
                sp_CalculateFilmPositions(camera, &filmP, &pixelPosition, 1);
                vec3 rayOrigin = camera->position;
                vec3 rayDirection = Normalize(filmP - camera->position);

                i32 object, triangle;
                f32 t = OraClosestHit(&s->scene, rayOrigin, rayDirection, &object, &triangle);

                // Cross-check against the reference's own entry point
                sp_Metrics scratch = {};
                sp_RayIntersectSceneResult check =
                    sp_RayIntersectScene(&s->scene, rayOrigin, rayDirection, &scratch);
                if (memcmp(&check.t, &t, sizeof(f32)) != 0)
                {
                    OraLog("oracle/_ref: harness t %.9g != sp_RayIntersectScene t %.9g at (%u,%u)",
                           t, check.t, x, y);
                    abort();
                }

                u32 index = x + y * width;
                if (triId) triId[index] = triangle;
                if (objId) objId[index] = object;
                if (tOut) tOut[index] = t;
                if (rayDir3)
                {
                    rayDir3[index * 3 + 0] = rayDirection.x;
                    rayDir3[index * 3 + 1] = rayDirection.y;
                    rayDir3[index * 3 + 2] = rayDirection.z;
                }
            }
        }
    };
    std::vector<std::thread> pool;
    for (u32 t = 1; t < threads; ++t) pool.emplace_back(worker, t);
    worker(0);
    for (auto &t : pool) t.join();
}

extern "C" void ora_intersect_rays(ora_Scene *s, uint32_t n, const float *origins3,
                                   const float *dirs3, float *out7, int32_t *triId,
                                   int32_t *objId, uint64_t *metrics)
{
    sp_Metrics m = {};
    for (u32 i = 0; i < n; ++i)
    {
        vec3 o = Vec3(origins3[i * 3], origins3[i * 3 + 1], origins3[i * 3 + 2]);
        vec3 d = Vec3(dirs3[i * 3], dirs3[i * 3 + 1], dirs3[i * 3 + 2]);
        sp_RayIntersectSceneResult r = sp_RayIntersectScene(&s->scene, o, d, &m);
        if (out7)
        {
            float *out = out7 + (size_t)i * 7;
            out[0] = r.t;
            memcpy(&out[1], &r.materialId, 4);
            out[2] = r.normal.x;
            out[3] = r.normal.y;
            out[4] = r.normal.z;
            out[5] = r.uv.x;
            out[6] = r.uv.y;
        }
        if (triId || objId)
        {
            i32 object, triangle;
            OraClosestHit(&s->scene, o, d, &object, &triangle);
            if (triId) triId[i] = triangle;
            if (objId) objId[i] = object;
        }
    }
    OraAddMetrics(metrics, m);
}

// ---- known-answer entry points --------------------------------------------------------------

extern "C" uint32_t ora_xorshift32(uint32_t *state)
{
    RandomNumberGenerator rng = {*state};
    u32 r = XorShift32(&rng);
    *state = rng.state;
    return r;
}

extern "C" float ora_random_unilateral(uint32_t *state)
{
    RandomNumberGenerator rng = {*state};
    f32 r = RandomUnilateral(&rng);
    *state = rng.state;
    return r;
}

extern "C" float ora_random_bilateral(uint32_t *state)
{
    RandomNumberGenerator rng = {*state};
    f32 r = RandomBilateral(&rng);
    *state = rng.state;
    return r;
}

extern "C" void ora_ray_triangle_mt(const float *o, const float *d, const float *a,
                                    const float *b, const float *c, float *out6)
{
    RayIntersectTriangleResult r = RayIntersectTriangleMT(Vec3(o[0], o[1], o[2]),
        Vec3(d[0], d[1], d[2]), Vec3(a[0], a[1], a[2]), Vec3(b[0], b[1], b[2]),
        Vec3(c[0], c[1], c[2]));
    out6[0] = r.t;
    out6[1] = r.uv.x;
    out6[2] = r.uv.y;
    out6[3] = r.normal.x;
    out6[4] = r.normal.y;
    out6[5] = r.normal.z;
}

extern "C" uint32_t ora_ray_aabb4(const float *boxMin12, const float *boxMax12, const float *o,
                                  const float *inv)
{
    vec3 mins[4], maxes[4];
    for (u32 i = 0; i < 4; ++i)
    {
        mins[i] = Vec3(boxMin12[i * 3], boxMin12[i * 3 + 1], boxMin12[i * 3 + 2]);
        maxes[i] = Vec3(boxMax12[i * 3], boxMax12[i * 3 + 1], boxMax12[i * 3 + 2]);
    }
    return simd_RayIntersectAabb4(mins, maxes, Vec3(o[0], o[1], o[2]),
                                  Vec3(inv[0], inv[1], inv[2]));
}

extern "C" float ora_ray_aabb_scalar(const float *mn, const float *mx, const float *o,
                                     const float *d)
{
    return RayIntersectAabb(Vec3(mn[0], mn[1], mn[2]), Vec3(mx[0], mx[1], mx[2]),
                            Vec3(o[0], o[1], o[2]), Vec3(d[0], d[1], d[2]));
}

extern "C" void ora_hemisphere(uint32_t *state, const float *n, float *out3)
{
    RandomNumberGenerator rng = {*state};
    vec3 v = RandomDirectionOnHemisphere(Vec3(n[0], n[1], n[2]), &rng);
    *state = rng.state;
    out3[0] = v.x;
    out3[1] = v.y;
    out3[2] = v.z;
}

extern "C" void ora_to_spherical(const float *v, float *out2)
{
    vec2 r = ToSphericalCoordinates(Vec3(v[0], v[1], v[2]));
    out2[0] = r.x;
    out2[1] = r.y;
}

extern "C" void ora_map_equirect(const float *sphere2, float *out2)
{
    vec2 r = MapToEquirectangular(Vec2(sphere2[0], sphere2[1]));
    out2[0] = r.x;
    out2[1] = r.y;
}

extern "C" void ora_spherical_to_cartesian(const float *sphere2, float *out3)
{
    vec3 r = MapSphericalToCartesianCoordinates(Vec2(sphere2[0], sphere2[1]));
    out3[0] = r.x;
    out3[1] = r.y;
    out3[2] = r.z;
}

extern "C" void ora_camera_fields(const float *p, const float *q, float filmDistance,
                                  uint32_t width, uint32_t height, float *out)
{
    ImagePlane plane = {};
    plane.width = width;
    plane.height = height;
    sp_Camera cam = {};
    quat rotation = {q[0], q[1], q[2], q[3]};
    sp_ConfigureCamera(&cam, &plane, Vec3(p[0], p[1], p[2]), rotation, filmDistance);
    vec3 v[5] = {cam.basis.right, cam.basis.up, cam.basis.forward, cam.position, cam.filmCenter};
    for (u32 i = 0; i < 5; ++i)
    {
        out[i * 3 + 0] = v[i].x;
        out[i * 3 + 1] = v[i].y;
        out[i * 3 + 2] = v[i].z;
    }
    out[15] = cam.halfPixelWidth;
    out[16] = cam.halfPixelHeight;
    out[17] = cam.halfFilmWidth;
    out[18] = cam.halfFilmHeight;
    out[19] = out[20] = out[21] = 0.0f;
}

extern "C" void ora_film_positions(ora_Scene *s, uint32_t n, const float *pixelPos2, float *out3)
{
    for (u32 i = 0; i < n; ++i)
    {
        vec2 p = Vec2(pixelPos2[i * 2], pixelPos2[i * 2 + 1]);
        vec3 f = {};
        sp_CalculateFilmPositions(&s->camera, &f, &p, 1);
        out3[i * 3 + 0] = f.x;
        out3[i * 3 + 1] = f.y;
        out3[i * 3 + 2] = f.z;
    }
}

extern "C" void ora_transform_aabb(const float *mn, const float *mx, const float *p,
                                   const float *q, const float *sc, float *out6)
{
    quat rotation = {q[0], q[1], q[2], q[3]};
    Aabb r = TransformAabb(Vec3(mn[0], mn[1], mn[2]), Vec3(mx[0], mx[1], mx[2]),
                           Vec3(p[0], p[1], p[2]), rotation, Vec3(sc[0], sc[1], sc[2]));
    out6[0] = r.min.x;
    out6[1] = r.min.y;
    out6[2] = r.min.z;
    out6[3] = r.max.x;
    out6[4] = r.max.y;
    out6[5] = r.max.z;
}

extern "C" void ora_radiance_for_path(ora_Scene *s, const float *path15, uint32_t n, float *out3)
{
    std::vector<sp_PathVertex> path(n ? n : 1);
    for (u32 i = 0; i < n; ++i)
    {
        const float *p = path15 + (size_t)i * 15;
        sp_PathVertex v = {};
        memcpy(&v.materialId, &p[0], 4);
        v.worldPosition = Vec3(p[1], p[2], p[3]);
        v.outgoingDir = Vec3(p[4], p[5], p[6]);
        v.incomingDir = Vec3(p[7], p[8], p[9]);
        v.normal = Vec3(p[10], p[11], p[12]);
        v.uv = Vec2(p[13], p[14]);
        path[i] = v;
    }
    vec3 r = ComputeRadianceForPath(path.data(), n, &s->materials);
    out3[0] = r.x;
    out3[1] = r.y;
    out3[2] = r.z;
}

extern "C" void ora_sample_nearest(const float *pixels, uint32_t w, uint32_t h, float u, float v,
                                   float *out4)
{
    HdrImage image = {(float *)pixels, w, h};
    vec4 r = SampleImageNearest(image, Vec2(u, v));
    memcpy(out4, r.data, 16);
}

extern "C" void ora_sample_bilinear(const float *pixels, uint32_t w, uint32_t h, float u,
                                    float v, float *out4)
{
    HdrImage image = {(float *)pixels, w, h};
    vec4 r = SampleImageBilinear(image, Vec2(u, v));
    memcpy(out4, r.data, 16);
}

extern "C" uint32_t ora_compute_tiles(uint32_t w, uint32_t h, uint32_t tw, uint32_t th,
                                      uint32_t *tiles, uint32_t maxTiles)
{
    return ComputeTiles(w, h, tw, th, (Tile *)tiles, maxTiles);
}

static void OraTreeWalk(bvh_Node *node, u32 depth, u32 *leafCount, u32 *internalCount,
                        u32 *minDepth, u32 *maxDepth, u32 *contained)
{
    if (node->children[0] == NULL)
    {
        (*leafCount)++;
        if (depth < *minDepth) *minDepth = depth;
        if (depth > *maxDepth) *maxDepth = depth;
        return;
    }
    (*internalCount)++;
    for (u32 i = 0; i < 4; ++i)
    {
        bvh_Node *child = node->children[i];
        if (child == NULL) continue;
        // Proper containment (aabb.h:60-74 uses && where || is meant, so check it here)
        for (u32 axis = 0; axis < 3; ++axis)
        {
            if (child->min.data[axis] < node->min.data[axis] ||
                child->max.data[axis] > node->max.data[axis])
            {
                *contained = 0;
            }
        }
        OraTreeWalk(child, depth + 1, leafCount, internalCount, minDepth, maxDepth, contained);
    }
}

extern "C" uint32_t ora_bvh_query(const float *aabbMin, const float *aabbMax, uint32_t n,
                                  const float *o, const float *d, uint32_t *leaves,
                                  uint32_t maxLeaves, uint32_t *error, uint32_t *aabbTests,
                                  float *rootBounds6)
{
    bvh_Tree tree = {};
    std::vector<u8> storage;
    if (n > 0)
    {
        size_t bytes = (size_t)n * 10 * sizeof(bvh_Node) +
                       2 * (size_t)n * sizeof(bvh_NodeDistSqPair) + 4096;
        storage.resize(bytes);
        MemoryArena arena;
        InitializeMemoryArena(&arena, storage.data(), bytes);
        tree = bvh_CreateTree(&arena, (vec3 *)aabbMin, (vec3 *)aabbMax, n);
    }
    std::vector<bvh_Node *> nodes(maxLeaves ? maxLeaves : 1);
    bvh_IntersectRayResult r = bvh_IntersectRay(&tree, Vec3(o[0], o[1], o[2]),
        Vec3(d[0], d[1], d[2]), nodes.data(), maxLeaves);
    for (u32 i = 0; i < r.count; ++i) leaves[i] = nodes[i]->leafIndex;
    if (error) *error = r.errorOccurred;
    if (aabbTests) *aabbTests = r.aabbTestCount;
    if (rootBounds6 && tree.root)
    {
        rootBounds6[0] = tree.root->min.x;
        rootBounds6[1] = tree.root->min.y;
        rootBounds6[2] = tree.root->min.z;
        rootBounds6[3] = tree.root->max.x;
        rootBounds6[4] = tree.root->max.y;
        rootBounds6[5] = tree.root->max.z;
    }
    return r.count;
}

// perf_tests/perf_tests.cpp:51-118 (TestBvh) and :212-305 (TestMeshMidphase) as calls: the reference's
// own functions, single-threaded, timed around the query loop only.
extern "C" double ora_perf_bvh(const float *aabbMin, const float *aabbMax, uint32_t n, uint32_t rays, const float *origins3,
                               const float *dirs3, uint32_t maxLeaves, uint32_t *countXorSum, double *buildSeconds)
{
    bvh_Tree tree = {};
    std::vector<u8> storage;
    auto b0 = std::chrono::steady_clock::now();
    if (n > 0)
    {
        size_t bytes = (size_t)n * 10 * sizeof(bvh_Node) + 2 * (size_t)n * sizeof(bvh_NodeDistSqPair) + 4096;
        storage.resize(bytes);
        MemoryArena arena;
        InitializeMemoryArena(&arena, storage.data(), bytes);
        tree = bvh_CreateTree(&arena, (vec3 *)aabbMin, (vec3 *)aabbMax, n);
    }
    if (buildSeconds) *buildSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - b0).count();
    std::vector<bvh_Node *> nodes(maxLeaves ? maxLeaves : 1);
    std::vector<uint32_t> counts(rays);
    auto t0 = std::chrono::steady_clock::now();
    for (uint32_t q = 0; q < rays; ++q)
    {
        bvh_IntersectRayResult r = bvh_IntersectRay(&tree, Vec3(origins3[q * 3], origins3[q * 3 + 1], origins3[q * 3 + 2]),
            Vec3(dirs3[q * 3], dirs3[q * 3 + 1], dirs3[q * 3 + 2]), nodes.data(), maxLeaves);
        // (the fingerprint is part of the timed loop on purpose: without a use of the result the
        // compiler may drop the query, and it is a few integer operations per reported leaf)
        uint32_t x = 0, sum = 0;
        for (u32 i = 0; i < r.count; ++i)
        {
            x ^= nodes[i]->leafIndex * 0x9E3779B1u;
            sum += nodes[i]->leafIndex;
        }
        countXorSum[q * 3 + 0] = r.count;
        countXorSum[q * 3 + 1] = x;
        countXorSum[q * 3 + 2] = sum;
    }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

extern "C" double ora_perf_mesh(ora_Scene *s, uint32_t mesh, uint32_t rays, const float *origins3, const float *dirs3,
                                float *t, int32_t *tri)
{
    sp_Metrics metrics = {};
    sp_Mesh m = s->meshes[mesh];
    auto t0 = std::chrono::steady_clock::now();
    for (uint32_t q = 0; q < rays; ++q)
    {
        sp_RayIntersectMeshResult r = sp_RayIntersectMesh(m, Vec3(origins3[q * 3], origins3[q * 3 + 1], origins3[q * 3 + 2]),
            Vec3(dirs3[q * 3], dirs3[q * 3 + 1], dirs3[q * 3 + 2]), &metrics);
        t[q] = r.triangleIntersection.t;
        if (tri) tri[q] = r.triangleIntersection.t >= 0.0f ? -2 : -1;
    }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

extern "C" void ora_mesh_tree_stats(ora_Scene *s, uint32_t mesh, uint32_t *out6)
{
    sp_Mesh *m = &s->meshes[mesh];
    u32 leafCount = 0, internalCount = 0, minDepth = 0xFFFFFFFFu, maxDepth = 0, contained = 1;
    OraTreeWalk(m->midphaseTree.root, 0, &leafCount, &internalCount, &minDepth, &maxDepth,
                &contained);
    u32 triangleCount = m->indexCount / 3;
    u32 reachable = 1;
    // bvh_FindLeafIndex is O(n) per query; fine for the fixture meshes
    for (u32 i = 0; i < triangleCount && triangleCount <= 20000; ++i)
    {
        if (!bvh_FindLeafIndex(m->midphaseTree.root, i))
        {
            reachable = 0;
            break;
        }
    }
    out6[0] = leafCount;
    out6[1] = internalCount;
    out6[2] = minDepth;
    out6[3] = maxDepth;
    out6[4] = reachable;
    out6[5] = contained;
}

// ---- environment pre-processing: the reference's bakers over a caller-supplied image --------
template <class CubeMap>
static void OraCopyFaces(const CubeMap &cube, uint32_t faceW, uint32_t faceH, float *out)
{
    for (uint32_t layer = 0; layer < 6; ++layer)
        memcpy(out + (size_t)layer * faceW * faceH * 4, cube.images[layer].pixels, (size_t)faceW * faceH * 16);
}

struct OraFaceArena
{
    MemoryArena arena;
    void *memory;
    OraFaceArena(uint32_t faceW, uint32_t faceH)
    {
        size_t bytes = (size_t)6 * faceW * faceH * 16 + 4096;
        memory = calloc(1, bytes);
        InitializeMemoryArena(&arena, memory, bytes);
    }
    ~OraFaceArena() { free(memory); }
};

extern "C" void ora_create_cube_map(const float *pixels, uint32_t w, uint32_t h, uint32_t faceW,
                                    uint32_t faceH, float *out)
{
    HdrImage env = {};
    env.pixels = (f32 *)pixels;
    env.width = w;
    env.height = h;
    OraFaceArena a(faceW, faceH);
    OraCopyFaces(ora_cubemap_uniform::CreateCubeMap(env, &a.arena, faceW, faceH), faceW, faceH, out);
}

extern "C" int ora_create_irradiance_cube_map(const float *pixels, uint32_t w, uint32_t h, uint32_t faceW,
                                              uint32_t faceH, uint32_t samplesPerPixel, uint32_t sampling,
                                              float sampleDelta, float *out)
{
    if (sampling == 0 && sampleDelta != 0.1f) return 0; // literal in cubemap.cpp:160
    HdrImage env = {};
    env.pixels = (f32 *)pixels;
    env.width = w;
    env.height = h;
    OraFaceArena a(faceW, faceH);
    if (sampling == 0)
        OraCopyFaces(ora_cubemap_uniform::CreateIrradianceCubeMap(env, &a.arena, faceW, faceH, samplesPerPixel),
                     faceW, faceH, out);
    else
        OraCopyFaces(ora_cubemap_random::CreateIrradianceCubeMap(env, &a.arena, faceW, faceH, samplesPerPixel),
                     faceW, faceH, out);
    return 1;
}
