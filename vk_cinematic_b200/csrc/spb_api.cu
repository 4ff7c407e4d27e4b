// spb_api.cu -- the C ABI of libspb200.so (include/sp_b200.h): scene capture, one-time upload,
// kernel launches, result copies.  Host logic only; the arithmetic of the path lives in
// spb_core.cuh and runs in the kernels of spb_kernels.cu.  There is no CPU implementation of any
// compute entry point in this library: without a usable CUDA device they abort.
#include <cuda_runtime.h>

#include <chrono>
#include <cmath>
#include <condition_variable>
#include <functional>
#include <thread>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <vector>

#include "spb_capture.h"
#include "spb_kernels.cuh"
#include "spb_cubemap.cuh"
#include "spb_strips.h"

using namespace spb;

struct PendingFrame; // (what sp_b200_RenderRowsBegin leaves for End; defined with the render entry points)

namespace {

// ---------------------------------------------------------------------------------------------
// logging / failure (reference convention: Assert -> LogMessage(file:line) -> abort,
// platform.h:18-22; tolerate a NULL sink)
sp_b200_LogFn g_log = nullptr;

void log_message(const char *fmt, ...)
{
    char buffer[1024];
    va_list args;
    va_start(args, fmt);
    vsnprintf(buffer, sizeof(buffer), fmt, args);
    va_end(args);
    if (g_log) g_log(buffer);
    else fprintf(stderr, "[sp_b200] %s\n", buffer);
}

#define SPB_ASSERT(X)                                                                   \
    do {                                                                                \
        if (!(X)) {                                                                     \
            log_message("%s:%d: Assertion failed!\n\t%s", __FILE__, __LINE__, #X);      \
            abort();                                                                    \
        }                                                                               \
    } while (0)

#define SPB_CUDA(call)                                                                  \
    do {                                                                                \
        cudaError_t err__ = (call);                                                     \
        if (err__ != cudaSuccess) {                                                     \
            log_message("%s:%d: CUDA error %s (%s) in %s -- libspb200 has no CPU path", \
                        __FILE__, __LINE__, cudaGetErrorName(err__),                    \
                        cudaGetErrorString(err__), #call);                              \
            abort();                                                                    \
        }                                                                               \
    } while (0)

// ---------------------------------------------------------------------------------------------
struct DeviceBuffer
{
    void *ptr = nullptr;
    size_t bytes = 0;
    // Stream-ordered form (the arrays of a scene): allocated on `allocStream` from the device's memory pool and
    // freed in the order of `freeStream` -- behind whatever frames in flight still read them -- instead of
    // with cudaMalloc / cudaFree, which wait for the whole device (sp_BuildSceneBroadphase between
    // sp_b200_RenderRowsBegin and End must not stall the frame in flight).
    const cudaStream_t *freeStream = nullptr; // (the owner's render stream variable: read when the buffer is freed)
    bool pooled = false;
    void ensure(size_t need)
    {
        if (need <= bytes) return;
        release();
        size_t grow = need + need / 4;
        SPB_CUDA(cudaMalloc(&ptr, grow));
        bytes = grow;
    }
    void ensure_pooled(size_t need, cudaStream_t allocStream, const cudaStream_t *freeOn)
    {
        release();
        SPB_CUDA(cudaMallocAsync(&ptr, need, allocStream));
        bytes = need;
        pooled = true;
        freeStream = freeOn;
    }
    void release()
    {
        if (ptr)
        {
            if (pooled) cudaFreeAsync(ptr, *freeStream);
            else cudaFree(ptr);
        }
        ptr = nullptr;
        bytes = 0;
        pooled = false;
    }
};

// What ONE frame in flight owns (sp_b200_RenderRowsBegin / End keep up to two): the events its launches
// are bracketed with, the pinned words its counters come back to, its device image.  Everything else a
// frame uses -- the wavefront working set, the device-side counters -- is shared: frames run one after
// the other on the render stream, and a frame's counters are copied out (in stream order) before the
// next frame's memset clears them.
struct FrameState
{
    cudaEvent_t evStart = nullptr, evKernel0 = nullptr, evKernel1 = nullptr, evEnd = nullptr;
    cudaEvent_t evRowsReady = nullptr, evCopyDone = nullptr;
    cudaEvent_t evSky0 = nullptr, evSky1 = nullptr; // around the sky kernels of a wavefront frame
    bool skyTimed = false;
    // CUDA events around every k_trace launch of the frame (sp_b200_Stats::traceMs)
    std::vector<cudaEvent_t> traceEvents;
    size_t traceEventsUsed = 0;
    uint32_t *coveredPinned = nullptr; // pinned words the asynchronous read-back of the coverage pass lands in
    bool coverSpeculated = false;      // render_wavefront ran on the cached covered-block count
    uint32_t specCovered = 0, specHash = 0; // ... namely these (the cache itself may belong to a later frame's strip by the time this one ends)
    DeviceBuffer image;
    // pinned landing zones of the frame's counters (device counters + per-tile-row cost; wave counters)
    unsigned long long *hostCounters = nullptr;
    uint32_t *hostWave = nullptr;
    size_t hostCountersCap = 0, hostWaveCap = 0; // elements
    bool pending = false;              // begun, not ended
};

struct DeviceScene
{
    uint32_t magic = 0x4E435353u; // "SSCN"
    DeviceBuffer nodes, tris, shade, objInv, objModel, objInfo, objTris, objBox;
    DScene d;
    uint64_t triangleCount = 0;
    uint64_t instancedTriangles = 0;
    uint32_t nodeCount = 0;
    uint32_t maxDepth = 0;
    size_t deviceBytes = 0;
    ~DeviceScene()
    {
        nodes.release(); tris.release(); shade.release();
        objInv.release(); objModel.release(); objInfo.release(); objTris.release(); objBox.release();
    }
};

struct TextureEntry
{
    DeviceBuffer buffer;
    const void *external = nullptr; // sp_b200_SetDeviceTexture: the caller's device copy (never freed here)
    uint32_t width = 0, height = 0;
    ~TextureEntry() { buffer.release(); }
};

#ifndef SPB_FUSE_MISS_DEFAULT
#define SPB_FUSE_MISS_DEFAULT 0 // (what -1, "by scene", means for single-object scenes; see render_wavefront)
#endif
#define SPB_EVICT_AUTO 0xFFFFFFFFu // (eviction threshold chosen by scene: evict_below())

struct Library
{
    std::recursive_mutex mutex;
    bool initialized = false;
    int device = 0;
    cudaStream_t stream = 0;
    sp_b200_Params params;
    bool statsEnabled = false;
    uint32_t pathsPerPass = 0; // 0: default (render_wavefront)
    bool skyCulling = true;    // sp_b200_SetSkyCulling
    bool sortBounceRays = true; // sp_b200_SetRaySorting
    bool primaryCandidates = true; // sp_b200_SetPrimaryCandidates
    uint32_t meshBuilder = 0;      // sp_b200_SetMeshBuilder
    int fuseMiss = -1;             // sp_b200_SetMissFusion: escaped rays shaded by the trace kernel where it retires them (-1: by scene)
    bool skyOneLookup = true;      // sp_b200_SetSkyCulling(2 = on with, 1 = on without the one-lookup path)
    uint32_t sortBounces = 1;   // bounces whose outgoing rays are direction-sorted (A/B knob)
    // refill thresholds of the trace kernel: primary rays, direction-sorted bounce rays, the rest.
    // 0 for the sorted class = measured choice between packet mode (1) and SPB_REFILL_THRESHOLD:
    // packets win when a tile's sorted rays are coherent (C3: one 8x4 block, 64 spp: -6 % frame
    // time), and lose on scenes of sub-pixel triangles (C5: +11 %).
    uint32_t refillThreshold[3] = {1, 0, 0}; // primary, sorted (0: measured), other (0: refill_other())
    // straggler eviction of the bounce traces (sp_b200_SetStragglerEviction): a packet of the direction-sorted
    // launch [0] / of the later launches [1] whose walking lanes drop below this many parks them in the
    // continuation buffer and a second launch walks the parked rays on, compacted; 0 = off
    uint32_t evictBelow[2] = {SPB_EVICT_AUTO, SPB_EVICT_AUTO}; // sorted, other launches (AUTO: by scene, evict_below())
    struct SortedTuner
    {
        unsigned long long signature = 0;
        double ms[2] = {0, 0}, rays[2] = {0, 0};
        int samples[2] = {0, 0};
        int choice = -1;       // index into candidates, -1 while measuring
        unsigned frames = 0;
        // this frame's measurements: candidate, pass (index into the wave counters), events
        struct Probe { int candidate; size_t ctrIndex; cudaEvent_t e0, e1; };
        std::vector<Probe> probes;
        std::vector<cudaEvent_t> pool;
    } tuner;
    FrameState frames[2];
    // what Begin leaves for End, per frame slot (kept with the Library, not with the calling thread: a host may
    // begin a frame on one thread and end it on another)
    std::shared_ptr<::PendingFrame> pendingFrames[2];
    FrameState *f = &frames[0]; // the frame the entry point at hand works on
    sp_b200_Stats lastStats;
    std::map<void *, std::shared_ptr<MeshAccel>> meshes;
    std::map<void *, std::unique_ptr<DeviceScene>> scenes;
    // objects beyond the reference's fixed table of 32 (sp_b200_AddObjectToScene), per sp_Scene
    struct ExtraObject { sp_Mesh mesh; u32 material; spbh::M4 model, invModel; float mn[3], mx[3]; };
    std::map<sp_Scene *, std::vector<ExtraObject>> extraObjects;
    std::map<const float *, std::unique_ptr<TextureEntry>> textures;
    // device allocations of flushed textures, reused by the next upload of the same size (a
    // flush forgets contents, not memory: cudaMalloc/cudaFree of 128 MiB per frame is slow)
    std::vector<std::unique_ptr<TextureEntry>> texturePool;
    std::unique_ptr<DeviceScene> emptyScene;
    DeviceBuffer counters, materials, scratchA, scratchB, scratchC;
    // wavefront working set (DESIGN.md "Data layout")
    DeviceBuffer wRays[2], wHitRec, wHitQ, wMissQ, wTerms, wRad, wCtr, wMask, wBlockList, wStage, wCand, wSkyList, wCont;
    // Copy engine side of a frame (DESIGN.md "End to end"): texture uploads run on copyStream and
    // the render stream waits for them only where the first kernel that reads a texture is
    // launched; finished rows go back to a pinned host image band by band on the same stream
    // while later bands render.
    cudaStream_t copyStream = nullptr;
    cudaStream_t uploadStream = nullptr; // scene uploads (upload_scene)
    cudaEvent_t evScene = nullptr;
    cudaEvent_t evTextures = nullptr, evOrder = nullptr;
    // Covered-block count of the last wavefront strip, keyed by everything the coverage pass depends
    // on: the next frame of the same strip is launched on that number without waiting for its own
    // coverage pass, and checked against it afterwards (render_wavefront).
    struct CoverCache
    {
        bool valid = false;
        uint64_t triangles = 0;
        uint32_t objects = 0;
        DCamera camera;
        uint32_t y0 = 0, y1 = 0, covered = 0, listHash = 0;
        bool skyCulling = true;
        // what the host derives from the list for the row copies: first block, last block of every band
        uint32_t firstBlock = 0, bandBlocks = 0;
        std::vector<uint32_t> bandLast;
    } coverCache;
    bool texturesPending = false; // an upload was issued on copyStream and nobody waited for it yet
    std::vector<cudaEvent_t> externalReady; // events of sp_b200_SetDeviceTexture copies nobody waited for yet
    bool overlapCopies = true;    // sp_b200_SetCopyOverlap

    Library()
    {
        memset(&lastStats, 0, sizeof(lastStats));
        params.samplesPerPixel = 1;
        params.bounceCount = 3;
        params.radianceClamp = 10.0f;
        params.envFilter = SP_B200_ENV_NEAREST;
        params.mathMode = SP_B200_MATH_F64_ROUNDED;
        params.cullByDistance = 1;
        params.tileWidth = 64;
        params.tileHeight = 64;
        params.renderMode = SP_B200_RENDER_WAVEFRONT;
        params.samplesPerPass = 0;
        params.triangleTest = SP_B200_TRIANGLE_MOLLER_TRUMBORE;
    }
};

// One Library per device.  The primary one serves every call made on the host's own threads (what
// round 1 had: a process-wide singleton); sp_b200_InitDevices adds one per further GPU, each owned
// by a worker thread that points `t_current` at it, so that the same entry points -- which all go
// through lib() -- render a strip on that device when the worker calls them.
thread_local Library *t_current = nullptr;

Library &primary_lib()
{
    static Library *instance = new Library(); // never destroyed: safe at process exit
    return *instance;
}

Library &lib() { return t_current ? *t_current : primary_lib(); }

// A further device of the process: its Library and the thread that owns it.  Jobs are run one at
// a time; run() returns when the job is done.
struct DeviceWorker
{
    Library *L = nullptr;
    std::thread thread;
    std::mutex m;
    std::condition_variable cv;
    std::function<void()> job;
    bool hasJob = false, done = true, quit = false;

    void start()
    {
        thread = std::thread([this] {
            t_current = L;
            for (;;)
            {
                std::function<void()> f;
                {
                    std::unique_lock<std::mutex> lock(m);
                    cv.wait(lock, [this] { return hasJob || quit; });
                    if (quit) return;
                    f = std::move(job);
                    hasJob = false;
                }
                f();
                {
                    std::lock_guard<std::mutex> lock(m);
                    done = true;
                }
                cv.notify_all();
            }
        });
    }
    void post(std::function<void()> f)
    {
        std::lock_guard<std::mutex> lock(m);
        job = std::move(f);
        hasJob = true;
        done = false;
        cv.notify_all();
    }
    void wait()
    {
        std::unique_lock<std::mutex> lock(m);
        cv.wait(lock, [this] { return done; });
    }
    void run(std::function<void()> f) { post(std::move(f)); wait(); }
    void stop()
    {
        {
            std::lock_guard<std::mutex> lock(m);
            quit = true;
        }
        cv.notify_all();
        if (thread.joinable()) thread.join();
    }
};

struct MultiDevice
{
    std::mutex mutex;                       // one multi-device frame at a time
    std::vector<int> devices;               // CUDA ordinals; [0] is the primary's
    std::vector<std::unique_ptr<DeviceWorker>> workers; // devices[1..]
    // host copies of the built scenes, for the further devices to upload from (keyed like
    // Library::scenes: sp_Scene::broadphaseTree.root)
    std::map<void *, std::shared_ptr<FlatScene>> flats;
    // strips of the last frame and what they cost (sp_b200_RenderFrame re-cuts from it)
    std::vector<uint32_t> bounds, nextBounds; // the cut of the last frame / the one the next frame will use
    uint32_t height = 0, quantum = 0;
    std::vector<sp_b200_Stats> stats;
    std::vector<double> seconds;
    bool active() const { return devices.size() > 1; }
};

MultiDevice &multi()
{
    static MultiDevice *instance = new MultiDevice();
    return *instance;
}

void ensure_init()
{
    Library &L = lib();
    if (L.initialized)
    {
        // the device is per-thread state of the CUDA runtime: a host thread that has not called
        // in before would otherwise launch on device 0
        int current = -1;
        if (cudaGetDevice(&current) == cudaSuccess && current != L.device) SPB_CUDA(cudaSetDevice(L.device));
        return;
    }
    int count = 0;
    cudaError_t err = cudaGetDeviceCount(&count);
    if (err != cudaSuccess || count == 0)
    {
        log_message("libspb200: no usable CUDA device (%s). This library is GPU-only; there is "
                    "no CPU fallback.", err != cudaSuccess ? cudaGetErrorString(err) : "0 devices");
        abort();
    }
    SPB_CUDA(cudaSetDevice(L.device));
    // (A/B knob) L2 fetch granularity: the shading kernels read scattered 16- and 32-byte records
    if (const char *g = getenv("SPB_B200_L2_FETCH"))
    {
        size_t before = 0, after = 0;
        cudaDeviceGetLimit(&before, cudaLimitMaxL2FetchGranularity);
        cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(g));
        cudaDeviceGetLimit(&after, cudaLimitMaxL2FetchGranularity);
        log_message("libspb200: L2 fetch granularity %zu -> %zu (%s)", before, after, cudaGetErrorString(e));
        cudaGetLastError();
    }
    if (const char *f = getenv("SPB_B200_FUSE_MISS")) L.fuseMiss = atoi(f); // (A/B and test runs; sp_b200_SetMissFusion)
    SPB_CUDA(cudaStreamCreateWithFlags(&L.copyStream, cudaStreamNonBlocking));
    SPB_CUDA(cudaStreamCreateWithFlags(&L.uploadStream, cudaStreamNonBlocking));
    SPB_CUDA(cudaEventCreateWithFlags(&L.evScene, cudaEventDisableTiming));
    {
        // the pool keeps what scenes free (a rebuild per frame is the reference's own pattern, main.cpp:1545-1552)
        cudaMemPool_t pool = nullptr;
        if (cudaDeviceGetDefaultMemPool(&pool, L.device) == cudaSuccess)
        {
            uint64_t keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    SPB_CUDA(cudaEventCreateWithFlags(&L.evTextures, cudaEventDisableTiming));
    SPB_CUDA(cudaEventCreateWithFlags(&L.evOrder, cudaEventDisableTiming));
    for (FrameState &F : L.frames)
    {
        SPB_CUDA(cudaEventCreate(&F.evStart));
        SPB_CUDA(cudaEventCreate(&F.evKernel0));
        SPB_CUDA(cudaEventCreate(&F.evKernel1));
        SPB_CUDA(cudaEventCreate(&F.evEnd));
        SPB_CUDA(cudaEventCreateWithFlags(&F.evRowsReady, cudaEventDisableTiming));
        SPB_CUDA(cudaEventCreate(&F.evCopyDone));
        SPB_CUDA(cudaEventCreate(&F.evSky0));
        SPB_CUDA(cudaEventCreate(&F.evSky1));
        SPB_CUDA(cudaHostAlloc((void **)&F.coveredPinned, 64, cudaHostAllocDefault));
        F.hostCountersCap = 8192;
        F.hostWaveCap = 16384;
        SPB_CUDA(cudaHostAlloc((void **)&F.hostCounters, F.hostCountersCap * 8, cudaHostAllocDefault));
        SPB_CUDA(cudaHostAlloc((void **)&F.hostWave, F.hostWaveCap * 4, cudaHostAllocDefault));
        F.pending = false;
    }
    L.f = &L.frames[0];
    L.initialized = true;
}

KernelConfig kernel_config()
{
    Library &L = lib();
    KernelConfig c;
    c.math = L.params.mathMode == SP_B200_MATH_FAST_F32 ? 1 : 0;
    c.envFilter = L.params.envFilter == SP_B200_ENV_BILINEAR ? 1 : 0;
    c.cull = L.params.cullByDistance ? 1 : 0;
    c.stats = L.statsEnabled ? 1 : 0;
    return c;
}

template <class T>
void upload(DeviceBuffer &dst, const std::vector<T> &src, cudaStream_t stream)
{
    size_t bytes = src.size() * sizeof(T);
    dst.ensure(bytes ? bytes : 16);
    if (bytes) SPB_CUDA(cudaMemcpyAsync(dst.ptr, src.data(), bytes, cudaMemcpyHostToDevice, stream));
}
// (a scene's array: pool memory, allocated and filled on the upload stream, freed in render-stream order)
template <class T>
void upload_pooled(DeviceBuffer &dst, const std::vector<T> &src, cudaStream_t uploadStream, const cudaStream_t *renderStream)
{
    size_t bytes = src.size() * sizeof(T);
    dst.ensure_pooled(bytes ? bytes : 16, uploadStream, renderStream);
    if (bytes) SPB_CUDA(cudaMemcpyAsync(dst.ptr, src.data(), bytes, cudaMemcpyHostToDevice, uploadStream));
}

std::unique_ptr<DeviceScene> upload_scene(const FlatScene &fs)
{
    Library &L = lib();
    auto ds = std::make_unique<DeviceScene>();
    // On the upload stream, which nothing else uses: a copy from pageable memory first waits for the stream it is
    // issued on, and the render stream may hold a frame in flight (sp_b200_RenderRowsBegin).  The copies have left
    // the host arrays when the calls return (staged); the render stream waits for their arrival on the device.
    upload_pooled(ds->nodes, fs.nodes, L.uploadStream, &L.stream);
    upload_pooled(ds->tris, fs.tris, L.uploadStream, &L.stream);
    upload_pooled(ds->shade, fs.shade, L.uploadStream, &L.stream);
    upload_pooled(ds->objInv, fs.objInv, L.uploadStream, &L.stream);
    upload_pooled(ds->objModel, fs.objModel, L.uploadStream, &L.stream);
    upload_pooled(ds->objInfo, fs.objInfo, L.uploadStream, &L.stream);
    std::vector<uint32_t> objTris = fs.objTris;
    if (objTris.empty()) objTris.push_back(0);
    upload_pooled(ds->objTris, objTris, L.uploadStream, &L.stream);
    upload_pooled(ds->objBox, fs.objBox, L.uploadStream, &L.stream);
    SPB_CUDA(cudaEventRecord(L.evScene, L.uploadStream));
    SPB_CUDA(cudaStreamWaitEvent(L.stream, L.evScene, 0));
    SPB_CUDA(cudaStreamWaitEvent(L.copyStream, L.evScene, 0));
    ds->d.nodes = (const v4f *)ds->nodes.ptr;
    ds->d.tris = (const v4f *)ds->tris.ptr;
    ds->d.shade = (const v4f *)ds->shade.ptr;
    ds->d.objInv = (const v4f *)ds->objInv.ptr;
    ds->d.objModel = (const v4f *)ds->objModel.ptr;
    ds->d.objInfo = (const v4u *)ds->objInfo.ptr;
    ds->d.objTris = (const uint32_t *)ds->objTris.ptr;
    ds->d.objBox = (const v4f *)ds->objBox.ptr;
    ds->d.tlasExtent = fs.tlasExtent;
    ds->d.tlasNodeCount = fs.tlasNodeCount;
    ds->instancedTriangles = fs.instancedTriangles;
    ds->d.triangleTest = 0;
    ds->d.tlasRoot = fs.tlasRoot;
    ds->d.objectCount = fs.objectCount;
    ds->triangleCount = fs.triangleCount;
    ds->nodeCount = (uint32_t)(fs.nodes.size() / 8);
    ds->maxDepth = fs.maxDepth;
    // the trace kernel pushes without a bound check (spb_core.cuh trav_push); two more entries are
    // its sentinels.  flatten_scene() has already kept the TLAS and every mesh tree inside their
    // shares (SPB_TLAS_STACK_LIMIT / SPB_MESH_STACK_LIMIT, which intersect_scene() relies on).
    if (fs.stackNeed + 4 > SPB_STACK_SIZE)
    {
        log_message("scene needs %u traversal stack entries, the kernels provide %u", fs.stackNeed + 4, SPB_STACK_SIZE);
        abort();
    }
    ds->deviceBytes = (fs.nodes.size() + fs.tris.size() + fs.shade.size() + fs.objInv.size() +
                       fs.objModel.size()) * sizeof(v4f) + fs.objInfo.size() * sizeof(v4u);
    return ds;
}

std::shared_ptr<MeshAccel> find_mesh(const sp_Mesh &mesh)
{
    Library &L = lib();
    if (!mesh.midphaseTree.root) return nullptr;
    auto it = L.meshes.find(mesh.midphaseTree.root);
    if (it == L.meshes.end())
    {
        log_message("sp_Mesh::midphaseTree was not built by sp_BuildMeshMidphase of libspb200");
        abort();
    }
    return it->second;
}

DeviceScene *find_scene(sp_Scene *scene)
{
    Library &L = lib();
    if (scene && scene->broadphaseTree.root)
    {
        auto it = L.scenes.find(scene->broadphaseTree.root);
        if (it == L.scenes.end())
        {
            // a further device (sp_b200_InitDevices) meets the scene for the first time: upload the
            // host copy sp_BuildSceneBroadphase kept
            std::shared_ptr<FlatScene> flat;
            if (t_current)
            {
                MultiDevice &M = multi();
                auto f = M.flats.find(scene->broadphaseTree.root);
                if (f != M.flats.end()) flat = f->second;
            }
            if (!flat)
            {
                log_message(t_current ? "sp_Scene was built before sp_b200_InitDevices: call sp_BuildSceneBroadphase again"
                                      : "sp_Scene::broadphaseTree was not built by sp_BuildSceneBroadphase of libspb200");
                abort();
            }
            it = L.scenes.emplace(scene->broadphaseTree.root, upload_scene(*flat)).first;
        }
        return it->second.get();
    }
    // a scene that was never built behaves like the reference's NULL root (bvh.cpp:218-229):
    // every ray misses
    if (!L.emptyScene)
    {
        std::vector<ObjectInstance> none;
        L.emptyScene = upload_scene(flatten_scene(none));
    }
    return L.emptyScene.get();
}

// the scene as a launch sees it: the uploaded arrays plus the per-call choices that live in DScene
DScene launch_scene(const DeviceScene *ds)
{
    DScene d = ds->d;
    d.triangleTest = lib().params.triangleTest == SP_B200_TRIANGLE_WATERTIGHT ? 1u : 0u;
    return d;
}

// Device copies of HdrImage pixel buffers, keyed by the host pointer (the reference aliases the
// caller's pixels; contents are assumed immutable until sp_b200_FlushTextureCache).
const v4f *device_texture(const HdrImage &image)
{
    Library &L = lib();
    if (!image.pixels || image.width == 0 || image.height == 0) return nullptr;
    auto it = L.textures.find(image.pixels);
    if (it != L.textures.end() && it->second->width == image.width && it->second->height == image.height)
        return (const v4f *)(it->second->external ? it->second->external : it->second->buffer.ptr);
    size_t bytes = (size_t)image.width * image.height * 16;
    std::unique_ptr<TextureEntry> entry;
    for (size_t i = 0; i < L.texturePool.size(); ++i)
        if (L.texturePool[i]->buffer.bytes >= bytes && L.texturePool[i]->buffer.bytes <= bytes + bytes / 2)
        {
            entry = std::move(L.texturePool[i]);
            L.texturePool.erase(L.texturePool.begin() + i);
            break;
        }
    if (!entry) entry = std::make_unique<TextureEntry>();
    entry->buffer.ensure(bytes);
    entry->width = image.width;
    entry->height = image.height;
    if (L.overlapCopies)
    {
        // no host wait: whoever launches the first kernel that reads a texture makes the render
        // stream wait for evTextures (wait_textures).  The buffer may have been read by kernels of
        // an earlier frame still in flight on the render stream: order the copy after them.
        SPB_CUDA(cudaEventRecord(L.evOrder, L.stream));
        SPB_CUDA(cudaStreamWaitEvent(L.copyStream, L.evOrder, 0));
        SPB_CUDA(cudaMemcpyAsync(entry->buffer.ptr, image.pixels, bytes, cudaMemcpyHostToDevice, L.copyStream));
        SPB_CUDA(cudaEventRecord(L.evTextures, L.copyStream));
        L.texturesPending = true;
    }
    else
    {
        SPB_CUDA(cudaMemcpyAsync(entry->buffer.ptr, image.pixels, bytes, cudaMemcpyHostToDevice, L.stream));
        SPB_CUDA(cudaStreamSynchronize(L.stream));
    }
    const v4f *p = (const v4f *)entry->buffer.ptr;
    L.textures[image.pixels] = std::move(entry);
    return p;
}

// The render stream goes on only after every texture upload issued so far has landed.
void wait_textures()
{
    Library &L = lib();
    for (cudaEvent_t e : L.externalReady) SPB_CUDA(cudaStreamWaitEvent(L.stream, e, 0));
    L.externalReady.clear();
    if (!L.texturesPending) return;
    SPB_CUDA(cudaStreamWaitEvent(L.stream, L.evTextures, 0));
    L.texturesPending = false;
}

// Uploads the material system (textures resolved) and returns the device pointer.  deferWait: the
// caller launches kernels that do not read textures first and calls wait_textures() itself.
const DMaterials *upload_materials(const sp_MaterialSystem *ms, size_t *bytesOut = nullptr, bool deferWait = false)
{
    Library &L = lib();
    static sp_MaterialSystem emptySystem; // zero-initialised
    if (!ms) ms = &emptySystem;
    const v4f *pixels[SPB_MAX_IMAGES] = {};
    for (uint32_t i = 0; i < ms->imageCount && i < SPB_MAX_IMAGES; ++i)
        pixels[i] = device_texture(ms->images[i]);
    DMaterials dm;
    convert_materials(ms, pixels, &dm);
    // an image that could not be uploaded (NULL pixels) must not be sampled
    for (uint32_t i = 0; i < dm.count; ++i)
    {
        if (dm.albedoImage[i] >= 0 && !dm.images[dm.albedoImage[i]].pixels) dm.albedoImage[i] = -1;
        if (dm.emissionImage[i] >= 0 && !dm.images[dm.emissionImage[i]].pixels) dm.emissionImage[i] = -1;
    }
    L.materials.ensure(sizeof(DMaterials));
    SPB_CUDA(cudaMemcpyAsync(L.materials.ptr, &dm, sizeof(dm), cudaMemcpyHostToDevice, L.stream));
    if (bytesOut) *bytesOut = sizeof(dm);
    if (!deferWait) wait_textures();
    return (const DMaterials *)L.materials.ptr;
}

unsigned long long *reset_counters(size_t extraSlots = 0)
{
    Library &L = lib();
    size_t bytes = (CTR_COUNT + extraSlots) * sizeof(unsigned long long);
    L.counters.ensure(bytes);
    SPB_CUDA(cudaMemsetAsync(L.counters.ptr, 0, bytes, L.stream));
    return (unsigned long long *)L.counters.ptr;
}

void add_metrics(sp_Metrics *metrics, const unsigned long long *c, float kernelMs)
{
    if (!metrics) return;
    // counters are incremented, CyclesElapsed is overwritten (simd_path_tracer.cpp:242,320,344);
    // "cycles" are device nanoseconds here
    metrics->values[sp_Metric_CyclesElapsed] = (u64)((double)kernelMs * 1.0e6);
    metrics->values[sp_Metric_PathsTraced] += c[CTR_PATHS];
    metrics->values[sp_Metric_RaysTraced] += c[CTR_RAYS];
    metrics->values[sp_Metric_RayHitCount] += c[CTR_HITS];
    metrics->values[sp_Metric_RayMissCount] += c[CTR_MISSES];
    metrics->values[sp_Metric_RayIntersectMesh_MidphaseAabbTestCount] += c[CTR_NODE_VISITS] * 4;
    metrics->values[sp_Metric_RayIntersectMesh_TestsPerformed] += c[CTR_OBJECT_TESTS];
}

void record_stats(const unsigned long long *c, float kernelMs, float totalMs)
{
    Library &L = lib();
    L.lastStats.rays = c[CTR_RAYS];
    L.lastStats.nodeVisits = c[CTR_NODE_VISITS];
    L.lastStats.triangleTests = c[CTR_TRIANGLE_TESTS];
    L.lastStats.objectTests = c[CTR_OBJECT_TESTS];
    L.lastStats.envClampedLookups = c[CTR_ENV_CLAMPED];
    L.lastStats.kernelMs = kernelMs;
    L.lastStats.totalMs = totalMs;
    L.lastStats.traceMs = kernelMs;
    L.lastStats.traceLaunches = 1;
    L.lastStats.tracedRays = c[CTR_RAYS];
}

spbh::M4 to_m4(const mat4 &m)
{
    spbh::M4 r;
    for (int c = 0; c < 4; ++c) r.c[c] = spbh::V4{m.columns[c].x, m.columns[c].y, m.columns[c].z, m.columns[c].w};
    return r;
}
mat4 from_m4(const spbh::M4 &m)
{
    mat4 r;
    for (int c = 0; c < 4; ++c) r.columns[c] = vec4{m.c[c].x, m.c[c].y, m.c[c].z, m.c[c].w};
    return r;
}

// device scene holding just this mesh as object 0 with an identity transform
DeviceScene *mesh_device_scene(const std::shared_ptr<MeshAccel> &accel, uint32_t smooth)
{
    if (!accel->deviceScene)
    {
        ObjectInstance ob;
        ob.mesh = accel;
        ob.material = 0;
        ob.smooth = smooth;
        ob.model = spbh::identity();
        ob.invModel = spbh::identity();
        for (int k = 0; k < 3; ++k)
        {
            ob.aabbMin[k] = accel->bvh.rootMin[k];
            ob.aabbMax[k] = accel->bvh.rootMax[k];
        }
        std::vector<ObjectInstance> one(1, ob);
        accel->deviceScene = upload_scene(flatten_scene(one)).release();
    }
    return (DeviceScene *)accel->deviceScene;
}

// Refill threshold of the unsorted bounce launches: a warp goes back to the queue when fewer lanes than this still
// have work.  Measured: 12 for one object (round 1), 16 for several (C5 with the four-class vote: 168.8 ms per
// frame against 171.8 with 12 and 173.0 with 8; profiles/r2/s16_ab_c5_vote_thresholds.txt).
// Straggler eviction thresholds (0: off) of the sorted (k = 0) and the other (k = 1) bounce launches.  Measured: one
// object 8 / off (C3 60.1 -> 58.8 ms, profiles/r2/s9_*); several objects 16 / 16 (C5 170.0 -> 165.5 ms against 8 / off,
// 167.8 with 12 / off, 171.7 with off / 12; profiles/r2/s17_*).
static uint32_t evict_below(int k, uint32_t objectCount)
{
    Library &L = lib();
    if (L.evictBelow[k] != SPB_EVICT_AUTO) return L.evictBelow[k];
    return objectCount > 1 ? 16u : (k == 0 ? 8u : 0u);
}
static uint32_t refill_other(uint32_t objectCount)
{
    Library &L = lib();
    if (L.refillThreshold[2]) return L.refillThreshold[2];
    return objectCount > 1 ? 16u : SPB_REFILL_THRESHOLD;
}

// Wavefront render of args' rectangle (the strip).  A coverage pass marks the 8x4 pixel blocks some
// triangle may project into; the pixels of all other blocks go to the sky kernel (one thread per
// pixel, no queues).  The covered blocks are rendered in bands of consecutive list entries, a band
// in passes of S samples per pixel (all of them when they fit), each pass a fixed sequence of
// kernels over device queues (spb_wavefront.cu).  One 4-byte read-back (the number of covered
// blocks) sizes the bands; nothing else returns to the host between kernels.  Band size and S are
// chosen so that one pass keeps about 32 Mi paths in flight.
// hostOut: pinned host image (full frame, same layout as ra.out) that finished rows are copied to on
// the copy stream while later bands render, or null (the caller copies).  Returns true when the rows
// [ra.y0, ra.y1) have been (asynchronously) copied to hostOut; evCopyDone then marks the last copy.
bool render_wavefront(const RenderArgs &ra, uint64_t instancedTriangles, std::vector<uint32_t> &countersOut, f32 *hostOut,
                      bool allowCached = true)
{
    Library &L = lib();
    const uint32_t spp = ra.spp, bounces = ra.bounces;
    KernelConfig cfg = kernel_config();
    const uint32_t width = ra.x1 - ra.x0, height = ra.y1 - ra.y0;
    const uint32_t blocksX = (width + 7) / 8, blocksY = (height + 3) / 4;
    const uint32_t blocks = blocksX * blocksY;
    // (statistics of this frame only, whatever path it takes below)
    L.f->traceEventsUsed = 0;
    L.tuner.probes.clear();
    countersOut.clear();

    WaveArgs a;
    memset(&a, 0, sizeof(a));
    a.scene = ra.scene;
    a.materials = ra.materials;
    a.camera = ra.camera;
    a.x0 = ra.x0; a.y0 = ra.y0; a.x1 = ra.x1; a.y1 = ra.y1;
    a.blocksX = blocksX; a.blocksY = blocksY;
    a.spp = spp;
    a.bounces = bounces;
    a.frame = ra.frame;
    a.clampValue = ra.clampValue;
    a.out = ra.out;
    a.stats = ra.counters;
    a.countStats = L.statsEnabled ? 1 : 0;
    a.tileRowCost = ra.tileRowCost;
    // (the second half of the caller's cost array: the sky kernels' class, see sp_b200_RenderRows)
    a.tileRowSky = ra.tileRowCost ? ra.tileRowCost + ((ra.y1 - 1) / ra.tileHeight - ra.y0 / ra.tileHeight + 1) : nullptr;
    a.tileHeight = ra.tileHeight;
    a.costRow0 = ra.y0 / (ra.tileHeight ? ra.tileHeight : 1);

    // coverage -> block mask + list (sky culling off: no triangle marks anything and the
    // "everything" flag is raised instead)
    // (sized for the whole image: a strip that grows at the next re-cut must not reallocate)
    const size_t imageBlocks = (size_t)blocksX * ((ra.camera.height + 3) / 4 + 1);
    L.wMask.ensure((imageBlocks > blocks ? imageBlocks : blocks) + 1);
    L.wBlockList.ensure((imageBlocks > blocks ? imageBlocks : blocks) * 4 + 8);
    uint32_t *listCount = (uint32_t *)L.wBlockList.ptr + blocks;
    a.blockMask = (const uint8_t *)L.wMask.ptr;
    launch_coverage(a, instancedTriangles, !L.skyCulling, (uint8_t *)L.wMask.ptr, (uint32_t *)L.wBlockList.ptr,
                    listCount, L.stream);
    // Sky pixels whose samples provably read one environment texel are settled with one lookup
    // (k_sky, spb_wavefront.cu); `spread` bounds how far a sample's direction can be from the
    // pixel centre's: (jitter + one rounding step of the pixel coordinate) x the angle of a pixel,
    // times 4, plus 1e-6 for the rounding of the normalisation.
    a.skyList = nullptr;
    a.skyDirectionSpread = 0.0f;
    if (L.skyOneLookup && ra.camera.width <= 65535u && ra.camera.height <= 65535u)
    {
        const DCamera &c = ra.camera;
        double fx = (double)c.filmCenter.x - c.position.x, fy = (double)c.filmCenter.y - c.position.y,
               fz = (double)c.filmCenter.z - c.position.z;
        double dist = sqrt(fx * fx + fy * fy + fz * fz);
        double jitter = fmax(fabs((double)c.halfPixelWidth), fabs((double)c.halfPixelHeight)) +
                        1.2e-7 * (double)(c.width > c.height ? c.width : c.height);
        double pixelAngle = dist > 0.0 ? 2.0 * fmax((double)c.halfFilmWidth / c.width, (double)c.halfFilmHeight / c.height) / dist : 1.0;
        double spread = 4.0 * jitter * pixelAngle + 1.0e-6;
        if (spread < 1.0e-4 && c.halfPixelWidth >= 0.0f && c.halfPixelHeight >= 0.0f)
        {
            L.wSkyList.ensure(((size_t)width * (ra.camera.height > height ? ra.camera.height : height) + 1) * 4);
            SPB_CUDA(cudaMemsetAsync(L.wSkyList.ptr, 0, 4, L.stream));
            a.skyList = (uint32_t *)L.wSkyList.ptr;
            a.skyDirectionSpread = (float)spread;
        }
    }
    // The one read-back of the frame: the number of covered blocks (it sizes the bands).  Only the
    // coverage pass is in front of it; the sky kernels -- the first to read the environment map --
    // are launched behind the first primary trace, which does not, so a texture upload in flight
    // on the copy stream has that long to land before anything waits for it.
    // A frame of the SAME strip, scene and camera as the last one has the same coverage: it is
    // launched on the cached count at once (no host wait in the middle of the frame: at 8 ms per
    // strip the round trip is 1 % of a GPU left idle) while the fresh count goes to a pinned word;
    // the caller compares the two when the frame is done and renders again, waiting this time, in the
    // case nobody has produced yet that they differ.  The key is what the coverage depends on as far as
    // the host can see cheaply (triangle and object counts, camera, rows): a scene rebuilt with the same
    // content keeps the cache, one whose triangles moved is caught by the comparison -- of the count and of
    // a fingerprint of the block list (k_list_blocks).
    Library::CoverCache &cc = L.coverCache;
    const bool sameCover = cc.valid && allowCached && cc.triangles == instancedTriangles && cc.objects == ra.scene.objectCount &&
                           cc.y0 == ra.y0 && cc.y1 == ra.y1 && cc.skyCulling == L.skyCulling &&
                           memcmp(&cc.camera, &ra.camera, sizeof(DCamera)) == 0;
    uint32_t covered = 0;
    L.f->coveredPinned[0] = 0xFFFFFFFFu;
    L.f->coveredPinned[1] = 0xFFFFFFFFu;
    SPB_CUDA(cudaMemcpyAsync(L.f->coveredPinned, listCount, 8, cudaMemcpyDeviceToHost, L.stream));
    if (sameCover)
    {
        covered = cc.covered;
        L.f->coverSpeculated = true;
        L.f->specCovered = cc.covered;
        L.f->specHash = cc.listHash;
    }
    else
    {
        SPB_CUDA(cudaStreamSynchronize(L.stream));
        covered = L.f->coveredPinned[0];
        L.f->coverSpeculated = false;
        cc.valid = true;
        cc.triangles = instancedTriangles;
        cc.objects = ra.scene.objectCount;
        cc.camera = ra.camera;
        cc.y0 = ra.y0;
        cc.y1 = ra.y1;
        cc.covered = covered;
        cc.listHash = L.f->coveredPinned[1];
        cc.skyCulling = L.skyCulling;
        cc.bandLast.clear();
        cc.bandBlocks = 0;
    }

    // rows -> host, band by band, on the copy stream (pinned destinations only: a pageable one
    // would block this thread inside cudaMemcpyAsync and stall the launches behind it)
    bool streamRows = false;
    if (hostOut && L.overlapCopies)
    {
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, hostOut) == cudaSuccess && attr.type == cudaMemoryTypeHost) streamRows = true;
        else cudaGetLastError();
    }
    // rows [rowBegin, rowEnd) are final once everything launched on the render stream so far is done
    auto copy_rows = [&](uint32_t rowBegin, uint32_t rowEnd) {
        if (!streamRows || rowEnd <= rowBegin) return;
        SPB_CUDA(cudaEventRecord(L.f->evRowsReady, L.stream));
        SPB_CUDA(cudaStreamWaitEvent(L.copyStream, L.f->evRowsReady, 0));
        const size_t offset = (size_t)rowBegin * ra.camera.width;
        SPB_CUDA(cudaMemcpyAsync(hostOut + offset * 4, ra.out + offset, (size_t)(rowEnd - rowBegin) * ra.camera.width * 16,
                                 cudaMemcpyDeviceToHost, L.copyStream));
        SPB_CUDA(cudaEventRecord(L.f->evCopyDone, L.copyStream));
    };
    uint32_t rowsCopied = ra.y0, rowsBottom = ra.y1; // rows outside [rowsCopied, rowsBottom) are on their way
    auto launch_sky_kernels = [&]() {
        wait_textures();
        SPB_CUDA(cudaEventRecord(L.f->evSky0, L.stream));
        launch_sky(cfg, a, L.stream);
        if (a.skyList) launch_sky_listed(cfg, a, L.stream);
        SPB_CUDA(cudaEventRecord(L.f->evSky1, L.stream));
        L.f->skyTimed = true;
    };
    if (covered == 0)
    {
        launch_sky_kernels();
        copy_rows(ra.y0, ra.y1);
        return streamRows;
    }

    const uint64_t targetItems = L.pathsPerPass ? L.pathsPerPass : (32u << 20);
    uint32_t S = L.params.samplesPerPass ? L.params.samplesPerPass : spp;
    if (S > spp) S = spp;
    if (32ull * S > targetItems) S = (uint32_t)(targetItems / 32);
    if (S < 1) S = 1;
    uint32_t bandBlocks = (uint32_t)(targetItems / (32ull * S));
    if (bandBlocks < 1) bandBlocks = 1;
    if (bandBlocks > covered) bandBlocks = covered;
    uint32_t bands = (covered + bandBlocks - 1) / bandBlocks;
    bandBlocks = (covered + bands - 1) / bands; // bands of equal size
    bands = (covered + bandBlocks - 1) / bandBlocks;
    const uint32_t passes = (spp + S - 1) / S;
    // last block of every band (the list is row-major): rows above the block row it lies in are
    // final once the band is done.  One more small read-back, the stream is idle anyway.
    std::vector<uint32_t> bandLast(bands, 0);
    uint32_t firstBlock = 0;
    if (streamRows && L.f->coverSpeculated && cc.bandBlocks == bandBlocks && cc.bandLast.size() == bands)
    {
        // (the list this frame will find is the one these were read from: its fingerprint is checked with the count)
        bandLast = cc.bandLast;
        firstBlock = cc.firstBlock;
    }
    else if (streamRows)
    {
        const uint32_t *list = (const uint32_t *)L.wBlockList.ptr;
        SPB_CUDA(cudaMemcpyAsync(&firstBlock, list, 4, cudaMemcpyDeviceToHost, L.stream));
        for (uint32_t b = 0; b < bands; ++b)
        {
            uint32_t last = (b + 1) * bandBlocks < covered ? (b + 1) * bandBlocks - 1 : covered - 1;
            SPB_CUDA(cudaMemcpyAsync(&bandLast[b], list + last, 4, cudaMemcpyDeviceToHost, L.stream));
        }
        SPB_CUDA(cudaStreamSynchronize(L.stream));
        if (L.f->coverSpeculated && (L.f->coveredPinned[0] != L.f->specCovered || L.f->coveredPinned[1] != L.f->specHash))
        {
            // (the stream has just been waited for: the fresh count is here already and differs from the cached one)
            cc.valid = false;
        }
        else
        {
            cc.bandLast = bandLast;
            cc.firstBlock = firstBlock;
            cc.bandBlocks = bandBlocks;
        }
    }
    SPB_ASSERT(32ull * bandBlocks * S < 0xFFFFFFFFull - SPB_QUEUE_SLACK);
    const uint32_t capacity = 32u * bandBlocks * S; // ray slots and path ids both fit

    // The working set is sized for the largest pass ANY strip of this image can need (the pass size
    // the whole image would use), not for this strip's: when the strips of a multi-GPU frame are
    // re-cut, a strip that grew must not stop for cudaFree / cudaMalloc in the middle of a frame
    // (measured: several ms, and a cost measurement the next cut cannot use).
    const uint64_t fullBlocks = (uint64_t)((ra.camera.width + 7) / 8) * ((ra.camera.height + 3) / 4);
    uint64_t allocBlocks = targetItems / (32ull * S);
    if (allocBlocks < 1) allocBlocks = 1;
    if (allocBlocks > fullBlocks) allocBlocks = fullBlocks;
    if (allocBlocks < bandBlocks) allocBlocks = bandBlocks;
    const size_t allocItems = (size_t)(32ull * S * allocBlocks);
    // queues and ray arrays carry slack for the partly filled chunks of the warps in flight
    const size_t slots = allocItems + SPB_QUEUE_SLACK;
    L.wRays[0].ensure(slots * 32);
    L.wRays[1].ensure(slots * 32);
    L.wHitRec.ensure(slots * 16);
    L.wHitQ.ensure(slots * 4);
    L.wMissQ.ensure(slots * 4);
    L.wTerms.ensure(allocItems * 32 * (bounces > 1 ? bounces - 1 : 1));
    L.wRad.ensure(allocItems * 16);
    // single-object scenes: per-pixel candidate triangles for the primary rays
    // (the padding of k_candidates covers a jitter of up to 1/100 pixel; sp_ConfigureCamera's is
    // 0.5 / width of a pixel, simd_path_tracer.cpp:17-18)
    const bool useCandidates = L.primaryCandidates && ra.scene.objectCount == 1 && ra.scene.tlasRoot != SPB_REF_EMPTY &&
                               ra.camera.halfPixelWidth <= 0.01f && ra.camera.halfPixelHeight <= 0.01f &&
                               ra.camera.halfPixelWidth >= 0.0f && ra.camera.halfPixelHeight >= 0.0f;
    if (useCandidates) L.wCand.ensure((size_t)allocBlocks * 32 * SPB_CAND_STRIDE * 4);
    a.sortPrimaryHits = (bounces > 1 && L.sortBounceRays) ? 1 : 0;
    if (a.sortPrimaryHits) L.wStage.ensure(slots * 32);
    a.stage = (v4f *)L.wStage.ptr;
    const size_t ctrWords = (size_t)bands * passes * bounces * WCTR_STRIDE;
    L.wCtr.ensure(ctrWords * 4);
    SPB_CUDA(cudaMemsetAsync(L.wCtr.ptr, 0, ctrWords * 4, L.stream));

    a.pathCapacity = capacity;
    a.rays[0] = (v4f *)L.wRays[0].ptr;
    a.rays[1] = (v4f *)L.wRays[1].ptr;
    a.hitRec = (v4f *)L.wHitRec.ptr;
    a.hitQ = (uint32_t *)L.wHitQ.ptr;
    a.missQ = (uint32_t *)L.wMissQ.ptr;
    a.pathTerms = (v4f *)L.wTerms.ptr;
    a.rad = (v4f *)L.wRad.ptr;
    {
        // Measured (profiles/r2/s24_* ... s26_*): C3 57.4 -> 56.3 ms per frame with it (k_trace +7.1 ms, k_shade_miss
        // -8.1), C5 165.5 -> 166.8; with the path's vertex terms prefetched when the ray starts: the same.  Off by
        // default: 1.8 % on one configuration, for 7 ms of DRAM-latency-bound shading inside the kernel whose
        // roofline the bench reports as traversal; a host that wants the frame time turns it on (mode 1).
        const int mode = L.fuseMiss >= 0 ? L.fuseMiss : (ra.scene.objectCount == 1 ? SPB_FUSE_MISS_DEFAULT : 0);
        a.fuseMiss = mode ? (1u | (cfg.math ? 2u : 0u) | (cfg.envFilter ? 4u : 0u)) : 0u;
    }

    // sorted-class refill threshold: fixed, chosen earlier, or being measured (alternating
    // candidates over the passes of the first frames, timed with events around the bounce-1 trace)
    static const uint32_t kCandidates[2] = {1u, SPB_REFILL_THRESHOLD};
    Library::SortedTuner &tn = L.tuner;
    const unsigned long long signature = instancedTriangles * 1000003ull + ra.scene.objectCount * 10007ull +
                                         (unsigned long long)spp * 101ull + S * 7ull + bounces + ((unsigned long long)width << 40);
    if (signature != tn.signature)
    {
        tn.signature = signature;
        tn.ms[0] = tn.ms[1] = tn.rays[0] = tn.rays[1] = 0.0;
        tn.samples[0] = tn.samples[1] = 0;
        tn.choice = -1;
        tn.frames = 0;
    }
    const bool tuning = a.sortPrimaryHits && L.refillThreshold[1] == 0 && tn.choice < 0;

    // continuation buffer of the evicting launches: half the ray slots (a launch that fills it stops evicting)
    const bool evicting = evict_below(0, ra.scene.objectCount) || evict_below(1, ra.scene.objectCount);
    a.cont = nullptr;
    a.contCapacity = 0;
    if (evicting)
    {
        const size_t records = allocItems / 2 + 1024;
        L.wCont.ensure(records * SPB_CONT_QUADS * 16);
        a.cont = (v4u *)L.wCont.ptr;
        a.contCapacity = (uint32_t)records;
    }

    auto timed_trace = [&](uint32_t bounce, bool primary, uint32_t evictBelow = 0) {
        if (L.f->traceEventsUsed + 2 > L.f->traceEvents.size())
            for (int k = 0; k < 64; ++k)
            {
                cudaEvent_t e;
                SPB_CUDA(cudaEventCreate(&e));
                L.f->traceEvents.push_back(e);
            }
        SPB_CUDA(cudaEventRecord(L.f->traceEvents[L.f->traceEventsUsed++], L.stream));
        if (evictBelow && !primary)
        {
            // the packets' stragglers are parked (EVICT), then walked on together (RESUME)
            const uint32_t keep = a.refillThreshold;
            a.refillThreshold = evictBelow;
            launch_wave_trace(cfg, a, bounce, SPB_TRACE_EVICT, L.stream);
            a.refillThreshold = refill_other(a.scene.objectCount);
            launch_wave_trace(cfg, a, bounce, SPB_TRACE_RESUME, L.stream);
            a.refillThreshold = keep;
        }
        else launch_wave_trace(cfg, a, bounce, primary ? SPB_TRACE_PRIMARY : SPB_TRACE_QUEUE, L.stream);
        SPB_CUDA(cudaEventRecord(L.f->traceEvents[L.f->traceEventsUsed++], L.stream));
    };

    uint32_t *ctr = (uint32_t *)L.wCtr.ptr;
    uint32_t passIndex = 0;
    for (uint32_t band = 0; band < bands; ++band)
    {
        const uint32_t first = band * bandBlocks;
        a.blockList = (const uint32_t *)L.wBlockList.ptr + first;
        a.bandBlocks = first + bandBlocks <= covered ? bandBlocks : covered - first;
        a.candidates = nullptr;
        if (useCandidates)
        {
            launch_candidates(a, (uint32_t *)L.wCand.ptr, L.stream);
            a.candidates = (const uint32_t *)L.wCand.ptr;
        }
        for (uint32_t pass = 0; pass < passes; ++pass)
        {
            a.firstSample = pass * S;
            a.samplesThisPass = spp - a.firstSample < S ? spp - a.firstSample : S;
            a.workItems = a.bandBlocks * 32u * a.samplesThisPass;
            a.ctr = ctr;
            const size_t ctrIndex = (size_t)(ctr - (uint32_t *)L.wCtr.ptr);
            ctr += (size_t)bounces * WCTR_STRIDE;
            a.refillThreshold = L.refillThreshold[0];
            timed_trace(0, true);
            if (band == 0 && pass == 0)
            {
                // sky pixels, behind the first primary trace (see above); the rows above the
                // first and below the last covered block row are sky only: off to the host
                launch_sky_kernels();
                if (streamRows)
                {
                    uint32_t top = ra.y0 + 4u * (firstBlock / blocksX);
                    uint32_t bottom = ra.y0 + 4u * (bandLast[bands - 1] / blocksX + 1u);
                    if (top > ra.y1) top = ra.y1;
                    if (bottom > ra.y1) bottom = ra.y1;
                    copy_rows(ra.y0, top);
                    copy_rows(bottom, ra.y1);
                    rowsCopied = top;
                    rowsBottom = bottom;
                }
            }
            for (uint32_t b = 0; b < bounces; ++b)
            {
                const bool sortedNext = b == 0 && a.sortPrimaryHits;
                if (a.sortPrimaryHits && b + 1 < bounces && b < L.sortBounces) launch_wave_shade_sorted(cfg, a, b, L.stream);
                else launch_wave_shade(cfg, a, b, L.stream);
                if (b + 1 >= bounces) break;
                int probe = -1;
                const uint32_t evictBelow = evict_below(sortedNext ? 0 : 1, ra.scene.objectCount);
                if (evictBelow)
                {
                    timed_trace(b + 1, false, evictBelow);
                    continue;
                }
                if (!sortedNext) a.refillThreshold = refill_other(a.scene.objectCount);
                else if (L.refillThreshold[1]) a.refillThreshold = L.refillThreshold[1];
                else if (tn.choice >= 0) a.refillThreshold = kCandidates[tn.choice];
                else
                {
                    probe = (int)((passIndex + tn.frames) & 1u);
                    a.refillThreshold = kCandidates[probe];
                }
                if (probe >= 0 && tuning && tn.probes.size() < 8)
                {
                    while (tn.pool.size() < 16)
                    {
                        cudaEvent_t e;
                        SPB_CUDA(cudaEventCreate(&e));
                        tn.pool.push_back(e);
                    }
                    Library::SortedTuner::Probe pr = {probe, ctrIndex, tn.pool[tn.probes.size() * 2], tn.pool[tn.probes.size() * 2 + 1]};
                    SPB_CUDA(cudaEventRecord(pr.e0, L.stream));
                    timed_trace(b + 1, false);
                    SPB_CUDA(cudaEventRecord(pr.e1, L.stream));
                    tn.probes.push_back(pr);
                }
                else timed_trace(b + 1, false);
            }
            passIndex++;
            launch_wave_accumulate(a, L.stream);
        }
        if (streamRows)
        {
            // the band's last block row may continue in the next band: stop above it
            uint32_t end = band + 1 < bands ? ra.y0 + 4u * (bandLast[band] / blocksX) : rowsBottom;
            if (end > rowsBottom) end = rowsBottom;
            if (end > rowsCopied)
            {
                copy_rows(rowsCopied, end);
                rowsCopied = end;
            }
        }
    }
    countersOut.resize(ctrWords);
    return streamRows;
}

// A scene handle is the address of the primary's DeviceScene, which the allocator may hand out
// again: the further devices must forget their copy the moment the primary's dies.
void forget_scene_on_devices(void *handle)
{
    MultiDevice &M = multi();
    M.flats.erase(handle);
    for (auto &w : M.workers)
        w->run([handle] {
            Library &L = lib();
            std::lock_guard<std::recursive_mutex> lock(L.mutex);
            if (L.initialized) cudaDeviceSynchronize();
            L.scenes.erase(handle);
        });
}

// everything one Library owns on its device (called on the thread that owns it)
void shutdown_library(Library &L)
{
    std::lock_guard<std::recursive_mutex> lock(L.mutex);
    if (L.initialized) cudaDeviceSynchronize();
    for (auto &m : L.meshes)
        if (m.second->deviceScene)
        {
            delete (DeviceScene *)m.second->deviceScene;
            m.second->deviceScene = nullptr;
        }
    L.meshes.clear();
    L.scenes.clear();
    L.textures.clear();
    L.texturePool.clear();
    L.emptyScene.reset();
    for (FrameState &F : L.frames) F.image.release();
    L.counters.release(); L.materials.release();
    L.scratchA.release(); L.scratchB.release(); L.scratchC.release();
    L.wRays[0].release(); L.wRays[1].release(); L.wHitRec.release();
    L.wHitQ.release(); L.wMissQ.release(); L.wTerms.release(); L.wRad.release(); L.wCtr.release();
    L.wMask.release(); L.wBlockList.release(); L.wStage.release(); L.wCand.release(); L.wSkyList.release(); L.wCont.release();
    for (FrameState &F : L.frames)
    {
        for (cudaEvent_t e : F.traceEvents) cudaEventDestroy(e);
        F.traceEvents.clear();
        F.traceEventsUsed = 0;
    }
    for (cudaEvent_t e : L.tuner.pool) cudaEventDestroy(e);
    L.tuner.pool.clear();
    L.tuner.probes.clear();
    L.tuner.signature = 0;
    if (L.initialized)
    {
        for (FrameState &F : L.frames)
        {
            cudaEventDestroy(F.evStart); cudaEventDestroy(F.evKernel0);
            cudaEventDestroy(F.evKernel1); cudaEventDestroy(F.evEnd);
            cudaEventDestroy(F.evRowsReady); cudaEventDestroy(F.evCopyDone);
            cudaEventDestroy(F.evSky0); cudaEventDestroy(F.evSky1);
            cudaFreeHost(F.coveredPinned); cudaFreeHost(F.hostCounters); cudaFreeHost(F.hostWave);
            F.coveredPinned = nullptr; F.hostCounters = nullptr; F.hostWave = nullptr;
            F.pending = false;
        }
        cudaEventDestroy(L.evTextures); cudaEventDestroy(L.evOrder);
        L.coverCache.valid = false;
        cudaStreamDestroy(L.copyStream);
        L.copyStream = nullptr;
        cudaStreamDestroy(L.uploadStream);
        L.uploadStream = nullptr;
        cudaEventDestroy(L.evScene);
        L.evScene = nullptr;
    }
    L.texturesPending = false;
    L.initialized = false;
}

void stop_devices()
{
    MultiDevice &M = multi();
    for (auto &w : M.workers)
    {
        w->run([] { shutdown_library(lib()); });
        w->stop();
        delete w->L;
    }
    M.workers.clear();
    M.devices.clear();
    M.flats.clear();
    M.bounds.clear();
    M.nextBounds.clear();
    M.stats.clear();
    M.seconds.clear();
}

} // namespace

// =============================================================================================
// library control

extern "C" int sp_b200_Init(int device)
{
    Library &L = lib();
    std::lock_guard<std::recursive_mutex> lock(L.mutex);
    if (L.initialized && device != L.device)
    {
        log_message("sp_b200_Init: already initialised on device %d", L.device);
        return 1;
    }
    L.device = device;
    ensure_init();
    return 0;
}

extern "C" void sp_b200_Shutdown(void)
{
    stop_devices();
    shutdown_library(primary_lib());
}

extern "C" int sp_b200_InitDeviceList(const i32 *deviceList, u32 listCount)
{
    MultiDevice &M = multi();
    std::lock_guard<std::mutex> lock(M.mutex);
    int count = 0;
    cudaError_t err = cudaGetDeviceCount(&count);
    if (err != cudaSuccess || count == 0)
    {
        log_message("sp_b200_InitDevices: no usable CUDA device (%s); libspb200 has no CPU path",
                    err != cudaSuccess ? cudaGetErrorString(err) : "0 devices");
        abort();
    }
    std::vector<int> wanted;
    for (u32 i = 0; i < listCount; ++i)
    {
        if (deviceList[i] < 0 || deviceList[i] >= count)
        {
            log_message("sp_b200_InitDevices: device %d of %d does not exist", deviceList[i], count);
            return -1;
        }
        wanted.push_back(deviceList[i]);
    }
    if (wanted.empty()) return -1;
    Library &P = primary_lib();
    if (M.active() || (P.initialized && P.device != wanted[0]))
    {
        log_message("sp_b200_InitDevices: devices were already chosen (call sp_b200_Shutdown first)");
        return -1;
    }
    if (sp_b200_Init(wanted[0]) != 0) return -1;
    M.devices = wanted;
    for (size_t i = 1; i < wanted.size(); ++i)
    {
        auto w = std::make_unique<DeviceWorker>();
        w->L = new Library();
        w->L->device = wanted[i];
        w->start();
        const int mine = wanted[i], first = wanted[0];
        w->run([mine, first] {
            ensure_init();
            // strips go to the primary's image with direct peer copies over NVLink where the
            // topology allows it (otherwise cudaMemcpyPeerAsync stages through the host)
            int can = 0;
            if (mine != first && cudaDeviceCanAccessPeer(&can, mine, first) == cudaSuccess && can)
                if (cudaDeviceEnablePeerAccess(first, 0) != cudaSuccess) cudaGetLastError();
        });
        M.workers.push_back(std::move(w));
    }
    for (size_t i = 1; i < wanted.size(); ++i)
    {
        int can = 0;
        if (wanted[i] != wanted[0] && cudaDeviceCanAccessPeer(&can, wanted[0], wanted[i]) == cudaSuccess && can)
            if (cudaDeviceEnablePeerAccess(wanted[i], 0) != cudaSuccess) cudaGetLastError();
    }
    M.bounds.clear();
    M.nextBounds.clear();
    return (int)wanted.size();
}

extern "C" int sp_b200_InitDevices(u32 deviceMask)
{
    i32 list[32];
    u32 n = 0;
    for (i32 d = 0; d < 32; ++d)
        if (deviceMask & (1u << d)) list[n++] = d;
    if (n == 0)
    {
        log_message("sp_b200_InitDevices: empty device mask");
        return -1;
    }
    return sp_b200_InitDeviceList(list, n);
}

extern "C" u32 sp_b200_DeviceCount(void)
{
    MultiDevice &M = multi();
    return M.active() ? (u32)M.devices.size() : 1u;
}

extern "C" void sp_b200_PartitionRows(u32 height, u32 quantum, u32 parts, const f64 *rowCost, u32 *bounds)
{
    partition_rows(height, quantum, parts, rowCost, bounds);
}

extern "C" void sp_b200_RowSeconds(u32 height, u32 quantum, u32 parts, const u32 *bounds, const f64 *units,
                                   const f64 *seconds, f64 *rowSeconds)
{
    std::vector<double> r = row_seconds(height, quantum, parts, bounds, units, seconds);
    for (size_t i = 0; i < r.size(); ++i) rowSeconds[i] = r[i];
}

extern "C" int sp_b200_GetDeviceStats(u32 index, sp_b200_Stats *stats, u32 *rowBegin, u32 *rowEnd)
{
    MultiDevice &M = multi();
    std::lock_guard<std::mutex> lock(M.mutex);
    if (index >= M.stats.size() || (size_t)index + 1 >= M.bounds.size()) return 1;
    if (stats) *stats = M.stats[index];
    if (rowBegin) *rowBegin = M.bounds[index];
    if (rowEnd) *rowEnd = M.bounds[index + 1];
    return 0;
}

extern "C" void sp_b200_SetLogCallback(sp_b200_LogFn fn) { g_log = fn; }
extern "C" void sp_b200_SetStream(void *cudaStream) { lib().stream = (cudaStream_t)cudaStream; }

extern "C" void sp_b200_DefaultParams(sp_b200_Params *params) { *params = Library().params; }

extern "C" void sp_b200_SetParams(const sp_b200_Params *params)
{
    SPB_ASSERT(params->samplesPerPixel >= 1);
    SPB_ASSERT(params->bounceCount >= 1 && params->bounceCount <= SPB_MAX_BOUNCES);
    SPB_ASSERT(params->tileWidth >= 1 && params->tileHeight >= 4 && params->tileHeight % 4 == 0);
    SPB_ASSERT(params->renderMode <= SP_B200_RENDER_PER_PIXEL);
    SPB_ASSERT(params->triangleTest <= SP_B200_TRIANGLE_WATERTIGHT);
    lib().params = *params;
}
extern "C" void sp_b200_GetParams(sp_b200_Params *params) { *params = lib().params; }
extern "C" void sp_b200_GetLastStats(sp_b200_Stats *stats) { *stats = lib().lastStats; }
extern "C" void sp_b200_EnableStats(int enable) { lib().statsEnabled = enable != 0; }
extern "C" u64 sp_b200_KernelLaunchCount(void) { return g_kernelLaunches; }

extern "C" void sp_b200_FlushTextureCache(void)
{
    Library &L = lib();
    std::lock_guard<std::recursive_mutex> lock(L.mutex);
    // The host pixels may change after this call: wait until every upload issued so far has left them.  Not for
    // the device -- a frame in flight (sp_b200_RenderRowsBegin) keeps reading its device copies, and whoever
    // reuses a pooled buffer orders its copy behind the render stream (device_texture).
    if (L.initialized)
    {
        if (L.overlapCopies) cudaEventSynchronize(L.evTextures);
        else cudaDeviceSynchronize();
    }
    for (auto &t : L.textures)
        if (!t.second->external) L.texturePool.push_back(std::move(t.second));
    L.textures.clear();
    L.externalReady.clear();
    while (L.texturePool.size() > SPB_MAX_IMAGES) L.texturePool.erase(L.texturePool.begin());
}

extern "C" void sp_b200_SetDeviceTexture(const f32 *hostPixels, const void *devicePixels, u32 width, u32 height, void *readyEvent)
{
    Library &L = lib();
    std::lock_guard<std::recursive_mutex> lock(L.mutex);
    ensure_init();
    SPB_ASSERT(hostPixels != nullptr);
    auto it = L.textures.find(hostPixels);
    if (it != L.textures.end())
    {
        if (!it->second->external) L.texturePool.push_back(std::move(it->second));
        L.textures.erase(it);
    }
    if (!devicePixels) return; // forget the association
    auto entry = std::make_unique<TextureEntry>();
    entry->external = devicePixels;
    entry->width = width;
    entry->height = height;
    L.textures[hostPixels] = std::move(entry);
    if (readyEvent) L.externalReady.push_back((cudaEvent_t)readyEvent);
}

extern "C" int sp_b200_AccumulateFrame(void *deviceAccum, const void *deviceFrame, const f32 *hostFrame, u32 pixelCount,
                                       u32 framesAccumulated, f32 *hostAccumOut)
{
    Library &L = lib();
    std::lock_guard<std::recursive_mutex> lock(L.mutex);
    ensure_init();
    SPB_ASSERT(deviceAccum != nullptr && (deviceFrame != nullptr || hostFrame != nullptr));
    if (pixelCount == 0) return 0;
    const size_t bytes = (size_t)pixelCount * 16;
    const v4f *frame = (const v4f *)deviceFrame;
    if (!frame)
    {
        L.scratchC.ensure(bytes);
        SPB_CUDA(cudaMemcpyAsync(L.scratchC.ptr, hostFrame, bytes, cudaMemcpyHostToDevice, L.stream));
        frame = (const v4f *)L.scratchC.ptr;
    }
    launch_accumulate_frame((v4f *)deviceAccum, frame, pixelCount, framesAccumulated, L.stream);
    SPB_CUDA(cudaGetLastError());
    if (hostAccumOut) SPB_CUDA(cudaMemcpyAsync(hostAccumOut, deviceAccum, bytes, cudaMemcpyDeviceToHost, L.stream));
    SPB_CUDA(cudaStreamSynchronize(L.stream));
    return 0;
}

extern "C" void sp_b200_SetPathsPerPass(u32 paths) { lib().pathsPerPass = paths; }
extern "C" void sp_b200_SetSkyCulling(int enable)
{
    lib().skyCulling = enable != 0;
    lib().skyOneLookup = enable != 1; // 1: coverage + sky kernel with the per-sample loop only
}
extern "C" void sp_b200_SetPrimaryCandidates(int enable) { lib().primaryCandidates = enable != 0; }
extern "C" void sp_b200_SetCopyOverlap(int enable) { lib().overlapCopies = enable != 0; }
extern "C" void sp_b200_SetRaySorting(int enable)
{
    lib().sortBounceRays = enable != 0;
    lib().sortBounces = enable > 1 ? (uint32_t)enable : 1u;
}
extern "C" void sp_b200_SetRefillThresholds(u32 primary, u32 sorted, u32 other)
{
    Library &L = lib();
    L.refillThreshold[0] = primary ? primary : 1;
    L.refillThreshold[1] = sorted; // 0: measured
    L.refillThreshold[2] = other; // 0: by scene (refill_other)
}

extern "C" void sp_b200_SetMissFusion(int mode) { lib().fuseMiss = mode < 0 ? -1 : (mode > 0 ? 1 : 0); }

extern "C" void sp_b200_SetStragglerEviction(u32 sorted, u32 other)
{
    Library &L = lib();
    L.evictBelow[0] = sorted > 32 ? 32 : sorted;
    L.evictBelow[1] = other > 32 ? 32 : other;
}

extern "C" void sp_b200_XorShift32Stream(u32 *state, u32 count, u32 *values)
{
    // XorShift32 (math_utils.h:184-196), `count` draws continuing *state: integer-only, host side
    u32 x = *state;
    for (u32 i = 0; i < count; ++i)
    {
        x ^= x << 13;
        x ^= x >> 17;
        x ^= x << 5;
        values[i] = x;
    }
    *state = x;
}

extern "C" u32 sp_b200_Seed(u32 pixelIndex, u32 sample, u32 frame)
{
    return stream_seed(pixelIndex, sample, frame); // integer-only: identical on host and device
}

// =============================================================================================
// scene construction (sp_scene.cpp)

extern "C" void sp_InitializeScene(sp_Scene *scene, MemoryArena *arena)
{
    // The reference sub-allocates 512 KiB for the broadphase (sp_scene.cpp:1-5); the library keeps
    // its structures in its own memory and only records an empty arena.
    (void)arena;
    scene->memoryArena.base = nullptr;
    scene->memoryArena.size = 0;
    scene->memoryArena.capacity = 0;
}

extern "C" sp_Mesh sp_CreateMesh(VertexPNT *vertices, u32 vertexCount, u32 *indices,
                                 u32 indexCount, b32 useSmoothShading)
{
    sp_Mesh result;
    memset(&result, 0, sizeof(result));
    result.vertices = vertices;
    result.vertexCount = vertexCount;
    result.indices = indices;
    result.indexCount = indexCount;
    result.useSmoothShading = useSmoothShading;
    return result;
}

// sp_b200_SetMeshBuilder(SP_B200_BUILDER_DEVICE_LBVH): the tree of the next sp_BuildMeshMidphase
// calls comes from the device builder (spb_lbvh.cu) + the host collapse; anything it cannot do
// (fewer than 8 triangles, a tree deeper than the traversal stack, a CUDA error) falls back to the
// host SAH builder and says so in sp_b200_BuildInfo.
static sp_b200_BuildInfo g_lastBuild;
static Bvh4 device_tree_builder(const float *aabbMin, const float *aabbMax, uint32_t count)
{
    Library &L = lib();
    g_lastBuild.builder = SP_B200_BUILDER_DEVICE_LBVH;
    g_lastBuild.fellBack = 1;
    g_lastBuild.deviceMs = 0.0f;
    Bvh4 out;
    if (count >= 8)
    {
        // the whole build on the device, the 4-wide collapse included; the host checks the finished tree
        // (SPB_B200_LBVH_HOST_COLLAPSE=1: round 1's path -- the binary tree comes back and the host collapses it)
        static const bool hostCollapse = getenv("SPB_B200_LBVH_HOST_COLLAPSE") && atoi(getenv("SPB_B200_LBVH_HOST_COLLAPSE")) != 0;
        float ms = 0.0f;
        if (!hostCollapse)
        {
            DeviceTree4 tree4;
            if (lbvh_build_bvh4_device(aabbMin, aabbMax, count, &tree4, &ms, L.stream) &&
                bvh4_adopt_device_tree(aabbMin, aabbMax, count, tree4, &out))
            {
                g_lastBuild.fellBack = 0;
                g_lastBuild.deviceMs = ms;
                return out;
            }
        }
        BinaryTree tree;
        if (lbvh_build_binary_device(aabbMin, aabbMax, count, &tree, &ms, L.stream) &&
            bvh4_from_binary(aabbMin, aabbMax, count, tree, &out))
        {
            g_lastBuild.fellBack = 0;
            g_lastBuild.deviceMs = ms;
            return out;
        }
    }
    return build_bvh4(aabbMin, aabbMax, count);
}

extern "C" void sp_b200_SetMeshBuilder(u32 builder)
{
    Library &L = lib();
    std::lock_guard<std::recursive_mutex> lock(L.mutex);
    SPB_ASSERT(builder == SP_B200_BUILDER_HOST_SAH || builder == SP_B200_BUILDER_DEVICE_LBVH);
    L.meshBuilder = builder;
}

extern "C" void sp_b200_GetLastBuildInfo(sp_b200_BuildInfo *info)
{
    Library &L = lib();
    std::lock_guard<std::recursive_mutex> lock(L.mutex);
    if (info) *info = g_lastBuild;
}

extern "C" void sp_BuildMeshMidphase(sp_Mesh *mesh, MemoryArena *arena, MemoryArena *tempArena)
{
    (void)arena;
    (void)tempArena;
    Library &L = lib();
    std::lock_guard<std::recursive_mutex> lock(L.mutex);
    SPB_ASSERT(mesh->indexCount % 3 == 0);
    for (u32 i = 0; i < mesh->indexCount; ++i) SPB_ASSERT(mesh->indices[i] < mesh->vertexCount);
    const bool onDevice = L.meshBuilder == SP_B200_BUILDER_DEVICE_LBVH;
    if (onDevice) ensure_init();
    memset(&g_lastBuild, 0, sizeof(g_lastBuild));
    auto t0 = std::chrono::steady_clock::now();
    std::shared_ptr<MeshAccel> accel =
        build_mesh_accel(mesh->vertices, mesh->vertexCount, mesh->indices, mesh->indexCount,
                         onDevice ? &device_tree_builder : nullptr);
    g_lastBuild.wallMs = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    g_lastBuild.triangleCount = mesh->indexCount / 3;
    g_lastBuild.nodeCount = (u32)accel->bvh.nodes.size();
    g_lastBuild.stackNeed = accel->bvh.stackNeed;
    g_lastBuild.maxDepth = accel->bvh.maxDepth;
    void *handle = accel.get();
    L.meshes[handle] = accel;
    mesh->midphaseTree.root = handle;
    mesh->midphaseTree.memoryPool.storage = nullptr;
    mesh->midphaseTree.memoryPool.objectSize = (u32)sizeof(Node4);
    mesh->midphaseTree.memoryPool.capacity = (u32)accel->bvh.nodes.size();
    mesh->midphaseTree.memoryPool.headIndex = 0xFFFFFFFFu;
}

extern "C" void sp_AddObjectToScene(sp_Scene *scene, sp_Mesh mesh, u32 material, vec3 position,
                                    quat orientation, vec3 scale)
{
    SPB_ASSERT(mesh.vertices != NULL);
    SPB_ASSERT(mesh.vertexCount > 0);
    spbh::M4 model, invModel;
    float mn[3], mx[3];
    compute_object_transform(nullptr, mesh.vertices, mesh.vertexCount, position, orientation,
                             scale, &model, &invModel, mn, mx);
    SPB_ASSERT(scene->objectCount < SP_SCENE_MAX_OBJECTS);
    u32 index = scene->objectCount++;
    scene->aabbMin[index] = vec3{mn[0], mn[1], mn[2]};
    scene->aabbMax[index] = vec3{mx[0], mx[1], mx[2]};
    scene->invModelMatrices[index] = from_m4(invModel);
    scene->modelMatrices[index] = from_m4(model);
    scene->meshes[index] = mesh;
    scene->materials[index] = material;
}

extern "C" u32 sp_b200_AddObjectToScene(sp_Scene *scene, sp_Mesh mesh, u32 material, vec3 position,
                                        quat orientation, vec3 scale)
{
    Library &L = lib();
    std::lock_guard<std::recursive_mutex> lock(L.mutex);
    if (scene->objectCount < SP_SCENE_MAX_OBJECTS)
    {
        // the fixed table is not full: anything kept beyond it belongs to an earlier build
        L.extraObjects.erase(scene);
        sp_AddObjectToScene(scene, mesh, material, position, orientation, scale);
        return scene->objectCount - 1;
    }
    SPB_ASSERT(mesh.vertices != NULL);
    SPB_ASSERT(mesh.vertexCount > 0);
    Library::ExtraObject ex;
    ex.mesh = mesh;
    ex.material = material;
    compute_object_transform(nullptr, mesh.vertices, mesh.vertexCount, position, orientation, scale,
                             &ex.model, &ex.invModel, ex.mn, ex.mx);
    std::vector<Library::ExtraObject> &list = L.extraObjects[scene];
    list.push_back(ex);
    return SP_SCENE_MAX_OBJECTS + (u32)list.size() - 1;
}

extern "C" u32 sp_b200_SceneObjectCount(sp_Scene *scene)
{
    Library &L = lib();
    std::lock_guard<std::recursive_mutex> lock(L.mutex);
    auto it = L.extraObjects.find(scene);
    u32 extra = (it != L.extraObjects.end() && scene->objectCount == SP_SCENE_MAX_OBJECTS) ? (u32)it->second.size() : 0;
    return scene->objectCount + extra;
}

extern "C" void sp_b200_ReleaseScene(sp_Scene *scene)
{
    Library &L = lib();
    std::lock_guard<std::recursive_mutex> lock(L.mutex);
    L.extraObjects.erase(scene);
    if (scene->broadphaseTree.root)
    {
        if (L.initialized) cudaDeviceSynchronize();
        L.scenes.erase(scene->broadphaseTree.root);
        forget_scene_on_devices(scene->broadphaseTree.root);
        scene->broadphaseTree.root = nullptr;
    }
}

extern "C" u64 sp_b200_SceneDeviceBytes(sp_Scene *scene)
{
    Library &L = lib();
    std::lock_guard<std::recursive_mutex> lock(L.mutex);
    if (!scene || !scene->broadphaseTree.root) return 0;
    auto it = L.scenes.find(scene->broadphaseTree.root);
    return it == L.scenes.end() ? 0 : (u64)it->second->deviceBytes;
}

extern "C" void sp_b200_ReleaseMesh(sp_Mesh *mesh)
{
    Library &L = lib();
    std::lock_guard<std::recursive_mutex> lock(L.mutex);
    auto it = L.meshes.find(mesh->midphaseTree.root);
    if (it != L.meshes.end())
    {
        if (it->second->deviceScene)
        {
            if (L.initialized) cudaDeviceSynchronize();
            delete (DeviceScene *)it->second->deviceScene;
            it->second->deviceScene = nullptr;
        }
        L.meshes.erase(it);
    }
    mesh->midphaseTree.root = nullptr;
}

extern "C" void sp_BuildSceneBroadphase(sp_Scene *scene)
{
    Library &L = lib();
    std::lock_guard<std::recursive_mutex> lock(L.mutex);
    ensure_init();
    // a rebuild of the same sp_Scene replaces its previous device copy (the reference app
    // rebuilds per render, main.cpp:1545-1552)
    if (scene->broadphaseTree.root && L.scenes.count(scene->broadphaseTree.root))
    {
        // (no wait: the old arrays are freed in render-stream order, behind the frames that still read them)
        L.scenes.erase(scene->broadphaseTree.root);
        forget_scene_on_devices(scene->broadphaseTree.root);
    }
    scene->broadphaseTree.root = nullptr;

    std::vector<ObjectInstance> objects(scene->objectCount);
    for (u32 i = 0; i < scene->objectCount; ++i)
    {
        ObjectInstance &ob = objects[i];
        ob.mesh = find_mesh(scene->meshes[i]);
        ob.material = scene->materials[i];
        ob.smooth = scene->meshes[i].useSmoothShading ? 1u : 0u;
        ob.model = to_m4(scene->modelMatrices[i]);
        ob.invModel = to_m4(scene->invModelMatrices[i]);
        ob.aabbMin[0] = scene->aabbMin[i].x; ob.aabbMin[1] = scene->aabbMin[i].y; ob.aabbMin[2] = scene->aabbMin[i].z;
        ob.aabbMax[0] = scene->aabbMax[i].x; ob.aabbMax[1] = scene->aabbMax[i].y; ob.aabbMax[2] = scene->aabbMax[i].z;
    }
    auto extra = L.extraObjects.find(scene);
    if (extra != L.extraObjects.end() && scene->objectCount == SP_SCENE_MAX_OBJECTS)
    {
        for (const Library::ExtraObject &ex : extra->second)
        {
            ObjectInstance ob;
            ob.mesh = find_mesh(ex.mesh);
            ob.material = ex.material;
            ob.smooth = ex.mesh.useSmoothShading ? 1u : 0u;
            ob.model = ex.model;
            ob.invModel = ex.invModel;
            for (int k = 0; k < 3; ++k) { ob.aabbMin[k] = ex.mn[k]; ob.aabbMax[k] = ex.mx[k]; }
            objects.push_back(ob);
        }
    }
    FlatScene fs = flatten_scene(objects);
    std::unique_ptr<DeviceScene> ds = upload_scene(fs);
    void *handle = ds.get();
    if (multi().active()) multi().flats[handle] = std::make_shared<FlatScene>(std::move(fs));
    scene->broadphaseTree.root = handle;
    scene->broadphaseTree.memoryPool.storage = nullptr;
    scene->broadphaseTree.memoryPool.objectSize = (u32)sizeof(Node4);
    scene->broadphaseTree.memoryPool.capacity = ds->nodeCount;
    scene->broadphaseTree.memoryPool.headIndex = 0xFFFFFFFFu;
    L.scenes[handle] = std::move(ds);
}

// =============================================================================================
// ray queries

extern "C" int sp_b200_RayIntersectSceneBatch(sp_Scene *scene, u32 count, const vec3 *rayOrigins,
                                              const vec3 *rayDirections,
                                              sp_RayIntersectSceneResult *results,
                                              i32 *triangleIndex, i32 *objectIndex,
                                              sp_Metrics *metrics)
{
    Library &L = lib();
    std::lock_guard<std::recursive_mutex> lock(L.mutex);
    ensure_init();
    if (count == 0) return 0;
    DeviceScene *ds = find_scene(scene);
    size_t rayBytes = (size_t)count * 12;
    L.scratchA.ensure(rayBytes);
    L.scratchB.ensure(rayBytes);
    L.scratchC.ensure((size_t)count * sizeof(HitRecord));
    unsigned long long *ctr = reset_counters();
    SPB_CUDA(cudaEventRecord(L.f->evStart, L.stream));
    SPB_CUDA(cudaMemcpyAsync(L.scratchA.ptr, rayOrigins, rayBytes, cudaMemcpyHostToDevice, L.stream));
    SPB_CUDA(cudaMemcpyAsync(L.scratchB.ptr, rayDirections, rayBytes, cudaMemcpyHostToDevice, L.stream));
    SPB_CUDA(cudaEventRecord(L.f->evKernel0, L.stream));
    launch_intersect_batch(kernel_config(), launch_scene(ds), count, (const float *)L.scratchA.ptr,
                           (const float *)L.scratchB.ptr, (HitRecord *)L.scratchC.ptr, ctr, L.stream);
    SPB_CUDA(cudaGetLastError());
    SPB_CUDA(cudaEventRecord(L.f->evKernel1, L.stream));
    std::vector<HitRecord> host(count);
    unsigned long long c[CTR_COUNT];
    SPB_CUDA(cudaMemcpyAsync(host.data(), L.scratchC.ptr, (size_t)count * sizeof(HitRecord), cudaMemcpyDeviceToHost, L.stream));
    SPB_CUDA(cudaMemcpyAsync(c, ctr, sizeof(c), cudaMemcpyDeviceToHost, L.stream));
    SPB_CUDA(cudaEventRecord(L.f->evEnd, L.stream));
    SPB_CUDA(cudaStreamSynchronize(L.stream));
    float kernelMs = 0, totalMs = 0;
    SPB_CUDA(cudaEventElapsedTime(&kernelMs, L.f->evKernel0, L.f->evKernel1));
    SPB_CUDA(cudaEventElapsedTime(&totalMs, L.f->evStart, L.f->evEnd));
    for (u32 i = 0; i < count; ++i)
    {
        if (results)
        {
            results[i].t = host[i].t;
            results[i].materialId = host[i].materialId;
            results[i].normal = vec3{host[i].nx, host[i].ny, host[i].nz};
            results[i].uv = vec2{host[i].u, host[i].v};
        }
        if (triangleIndex) triangleIndex[i] = host[i].triangle;
        if (objectIndex) objectIndex[i] = host[i].object;
    }
    if (metrics)
    {
        // sp_RayIntersectScene itself only touches the intersection counters
        metrics->values[sp_Metric_RayIntersectMesh_MidphaseAabbTestCount] += c[CTR_NODE_VISITS] * 4;
        metrics->values[sp_Metric_RayIntersectMesh_TestsPerformed] += c[CTR_OBJECT_TESTS];
        metrics->values[sp_Metric_CyclesElapsed_RayIntersectScene] += (u64)((double)kernelMs * 1.0e6);
    }
    record_stats(c, kernelMs, totalMs);
    return 0;
}

extern "C" sp_RayIntersectSceneResult sp_RayIntersectScene(sp_Scene *scene, vec3 rayOrigin,
                                                           vec3 rayDirection, sp_Metrics *metrics)
{
    sp_RayIntersectSceneResult result;
    memset(&result, 0, sizeof(result));
    sp_b200_RayIntersectSceneBatch(scene, 1, &rayOrigin, &rayDirection, &result, nullptr, nullptr, metrics);
    return result;
}

extern "C" sp_RayIntersectMeshResult sp_RayIntersectMesh(sp_Mesh mesh, vec3 rayOrigin,
                                                         vec3 rayDirection, sp_Metrics *metrics)
{
    Library &L = lib();
    std::lock_guard<std::recursive_mutex> lock(L.mutex);
    ensure_init();
    sp_RayIntersectMeshResult result;
    memset(&result, 0, sizeof(result));
    result.triangleIntersection.t = -1.0f;
    std::shared_ptr<MeshAccel> accel = find_mesh(mesh);
    if (!accel) return result; // midphase never built: NULL root, no intersections
    DeviceScene *ds = mesh_device_scene(accel, mesh.useSmoothShading);
    L.scratchA.ensure(32);
    L.scratchC.ensure(sizeof(HitRecord));
    float ray[6] = {rayOrigin.x, rayOrigin.y, rayOrigin.z, rayDirection.x, rayDirection.y, rayDirection.z};
    SPB_CUDA(cudaMemcpyAsync(L.scratchA.ptr, ray, sizeof(ray), cudaMemcpyHostToDevice, L.stream));
    launch_intersect_mesh(kernel_config(), launch_scene(ds), mesh.useSmoothShading ? 1u : 0u,
                          (const float *)L.scratchA.ptr, (const float *)L.scratchA.ptr + 3,
                          (HitRecord *)L.scratchC.ptr, L.stream);
    SPB_CUDA(cudaGetLastError());
    HitRecord h;
    SPB_CUDA(cudaMemcpyAsync(&h, L.scratchC.ptr, sizeof(h), cudaMemcpyDeviceToHost, L.stream));
    SPB_CUDA(cudaStreamSynchronize(L.stream));
    result.triangleIntersection.t = h.t;
    result.triangleIntersection.uv = vec2{h.u, h.v};
    result.triangleIntersection.normal = vec3{h.nx, h.ny, h.nz};
    (void)metrics;
    return result;
}

extern "C" int sp_b200_RayIntersectAabb4Batch(u32 count, const f32 *boxMin, const f32 *boxMax, const vec3 *rayOrigins,
                                              const vec3 *invRayDirections, u32 *masks, f32 *tnear)
{
    Library &L = lib();
    std::lock_guard<std::recursive_mutex> lock(L.mutex);
    ensure_init();
    if (count == 0) return 0;
    const size_t boxBytes = (size_t)count * 12 * 4, vecBytes = (size_t)count * 12;
    L.scratchA.ensure(boxBytes * 2 + vecBytes * 2);
    L.scratchB.ensure((size_t)count * (3 + 4) * 4);
    char *in = (char *)L.scratchA.ptr;
    SPB_CUDA(cudaMemcpyAsync(in, boxMin, boxBytes, cudaMemcpyHostToDevice, L.stream));
    SPB_CUDA(cudaMemcpyAsync(in + boxBytes, boxMax, boxBytes, cudaMemcpyHostToDevice, L.stream));
    SPB_CUDA(cudaMemcpyAsync(in + boxBytes * 2, rayOrigins, vecBytes, cudaMemcpyHostToDevice, L.stream));
    SPB_CUDA(cudaMemcpyAsync(in + boxBytes * 2 + vecBytes, invRayDirections, vecBytes, cudaMemcpyHostToDevice, L.stream));
    uint32_t *dMasks = (uint32_t *)L.scratchB.ptr;
    float *dNear = (float *)(dMasks + (size_t)count * 3);
    launch_slab_kat(count, (const float *)in, (const float *)(in + boxBytes), (const float *)(in + boxBytes * 2),
                    (const float *)(in + boxBytes * 2 + vecBytes), dMasks, dNear, L.stream);
    SPB_CUDA(cudaGetLastError());
    SPB_CUDA(cudaMemcpyAsync(masks, dMasks, (size_t)count * 3 * 4, cudaMemcpyDeviceToHost, L.stream));
    if (tnear) SPB_CUDA(cudaMemcpyAsync(tnear, dNear, (size_t)count * 4 * 4, cudaMemcpyDeviceToHost, L.stream));
    SPB_CUDA(cudaStreamSynchronize(L.stream));
    return 0;
}

extern "C" u32 sp_b200_MeshIntersectedLeaves(sp_Mesh mesh, vec3 rayOrigin, vec3 rayDirection,
                                             u32 *leafIndices, u32 maxIntersections,
                                             b32 *errorOccurred)
{
    Library &L = lib();
    std::lock_guard<std::recursive_mutex> lock(L.mutex);
    ensure_init();
    if (errorOccurred) *errorOccurred = 0;
    std::shared_ptr<MeshAccel> accel = find_mesh(mesh);
    if (!accel) return 0;
    DeviceScene *ds = mesh_device_scene(accel, mesh.useSmoothShading);
    L.scratchA.ensure(32);
    L.scratchB.ensure((size_t)(maxIntersections ? maxIntersections : 1) * 4);
    L.scratchC.ensure(16);
    float ray[6] = {rayOrigin.x, rayOrigin.y, rayOrigin.z, rayDirection.x, rayDirection.y, rayDirection.z};
    SPB_CUDA(cudaMemcpyAsync(L.scratchA.ptr, ray, sizeof(ray), cudaMemcpyHostToDevice, L.stream));
    launch_collect_leaves(launch_scene(ds), (const float *)L.scratchA.ptr, (const float *)L.scratchA.ptr + 3,
                          (uint32_t *)L.scratchB.ptr, maxIntersections, (uint32_t *)L.scratchC.ptr, L.stream);
    SPB_CUDA(cudaGetLastError());
    uint32_t ce[2];
    SPB_CUDA(cudaMemcpyAsync(ce, L.scratchC.ptr, sizeof(ce), cudaMemcpyDeviceToHost, L.stream));
    SPB_CUDA(cudaStreamSynchronize(L.stream));
    if (ce[0] && leafIndices)
        SPB_CUDA(cudaMemcpy(leafIndices, L.scratchB.ptr, (size_t)ce[0] * 4, cudaMemcpyDeviceToHost));
    if (errorOccurred) *errorOccurred = ce[1];
    return ce[0];
}

// Rays stay on the device between the two batched queries when the caller passes the same arrays
// again (perf tests time the query, not the upload): kernelMs is the CUDA-event time of the kernel.
static int mesh_batch(sp_Mesh mesh, u32 count, const vec3 *origins, const vec3 *dirs, int what, u32 *leafOut3, f32 *tOut,
                      i32 *triOut, f32 *kernelMs)
{
    Library &L = lib();
    std::lock_guard<std::recursive_mutex> lock(L.mutex);
    ensure_init();
    if (kernelMs) *kernelMs = 0.0f;
    if (count == 0) return 0;
    std::shared_ptr<MeshAccel> accel = find_mesh(mesh);
    if (!accel) return 1;
    DeviceScene *ds = mesh_device_scene(accel, mesh.useSmoothShading);
    const size_t vecBytes = (size_t)count * 12;
    L.scratchA.ensure(vecBytes * 2);
    L.scratchB.ensure((size_t)count * 12);
    char *in = (char *)L.scratchA.ptr;
    SPB_CUDA(cudaMemcpyAsync(in, origins, vecBytes, cudaMemcpyHostToDevice, L.stream));
    SPB_CUDA(cudaMemcpyAsync(in + vecBytes, dirs, vecBytes, cudaMemcpyHostToDevice, L.stream));
    SPB_CUDA(cudaEventRecord(L.f->evKernel0, L.stream));
    if (what == 0)
        launch_collect_leaves_batch(launch_scene(ds), count, (const float *)in, (const float *)(in + vecBytes), (uint32_t *)L.scratchB.ptr, L.stream);
    else
        launch_intersect_mesh_batch(kernel_config(), launch_scene(ds), count, (const float *)in, (const float *)(in + vecBytes),
                                    (float *)L.scratchB.ptr, (int32_t *)((float *)L.scratchB.ptr + count), L.stream);
    SPB_CUDA(cudaGetLastError());
    SPB_CUDA(cudaEventRecord(L.f->evKernel1, L.stream));
    if (what == 0)
        SPB_CUDA(cudaMemcpyAsync(leafOut3, L.scratchB.ptr, (size_t)count * 12, cudaMemcpyDeviceToHost, L.stream));
    else
    {
        if (tOut) SPB_CUDA(cudaMemcpyAsync(tOut, L.scratchB.ptr, (size_t)count * 4, cudaMemcpyDeviceToHost, L.stream));
        if (triOut) SPB_CUDA(cudaMemcpyAsync(triOut, (float *)L.scratchB.ptr + count, (size_t)count * 4, cudaMemcpyDeviceToHost, L.stream));
    }
    SPB_CUDA(cudaStreamSynchronize(L.stream));
    float ms = 0.0f;
    SPB_CUDA(cudaEventElapsedTime(&ms, L.f->evKernel0, L.f->evKernel1));
    if (kernelMs) *kernelMs = ms;
    return 0;
}

extern "C" int sp_b200_MeshIntersectedLeavesBatch(sp_Mesh mesh, u32 count, const vec3 *rayOrigins, const vec3 *rayDirections,
                                                  u32 *countXorSum, f32 *kernelMs)
{
    return mesh_batch(mesh, count, rayOrigins, rayDirections, 0, countXorSum, nullptr, nullptr, kernelMs);
}

extern "C" int sp_b200_RayIntersectMeshBatch(sp_Mesh mesh, u32 count, const vec3 *rayOrigins, const vec3 *rayDirections, f32 *t,
                                             i32 *triangleIndex, f32 *kernelMs)
{
    return mesh_batch(mesh, count, rayOrigins, rayDirections, 1, nullptr, t, triangleIndex, kernelMs);
}

extern "C" void sp_b200_MeshTreeInfo(sp_Mesh mesh, sp_b200_TreeInfo *info)
{
    Library &L = lib();
    std::lock_guard<std::recursive_mutex> lock(L.mutex);
    memset(info, 0, sizeof(*info));
    std::shared_ptr<MeshAccel> accel = find_mesh(mesh);
    if (!accel) return;
    const Bvh4 &bvh = accel->bvh;
    info->leafCount = (u32)bvh.slotPrim.size();
    info->nodeCount = (u32)bvh.nodes.size();
    info->maxDepth = bvh.maxDepth;
    info->rootMin = vec3{bvh.rootMin[0], bvh.rootMin[1], bvh.rootMin[2]};
    info->rootMax = vec3{bvh.rootMax[0], bvh.rootMax[1], bvh.rootMax[2]};
    // reachability and containment by walking the tree (cf. test_bvh.cpp:74-163)
    std::vector<uint8_t> seen(accel->triangleCount, 0);
    bool contained = true;
    std::vector<uint32_t> work;
    if (!bvh.nodes.empty()) work.push_back(0);
    while (!work.empty())
    {
        uint32_t ni = work.back();
        work.pop_back();
        const Node4 &n = bvh.nodes[ni];
        for (int k = 0; k < 4; ++k)
        {
            uint32_t r = n.ref[k];
            if (r == SPB_REF_EMPTY) continue;
            if (r & SPB_REF_LEAF)
            {
                uint32_t prim = bvh.slotPrim[r & ~SPB_REF_LEAF];
                if (prim < seen.size()) seen[prim] = 1;
                continue;
            }
            const Node4 &c = bvh.nodes[r];
            for (int j = 0; j < 4; ++j)
            {
                if (c.ref[j] == SPB_REF_EMPTY) continue;
                for (int a = 0; a < 3; ++a)
                    if (c.bmin[a][j] < n.bmin[a][k] || c.bmax[a][j] > n.bmax[a][k]) contained = false;
            }
            work.push_back(r);
        }
    }
    bool all = true;
    for (uint8_t s : seen) all = all && s;
    info->allLeavesReachable = all;
    info->parentsContainChildren = contained;
}

// =============================================================================================
// materials (sp_material_system.cpp) -- table maintenance is plain host code, evaluation is GPU

extern "C" b32 sp_RegisterMaterial(sp_MaterialSystem *materialSystem, sp_Material material, u32 id)
{
    if (materialSystem->count >= SP_MAX_MATERIALS) return 0;
    u32 index = materialSystem->count++;
    materialSystem->keys[index] = id;
    materialSystem->materials[index] = material;
    return 1;
}

extern "C" sp_Material *sp_FindMaterialById(sp_MaterialSystem *materialSystem, u32 id)
{
    for (u32 i = 0; i < materialSystem->count; ++i)
        if (materialSystem->keys[i] == id) return materialSystem->materials + i;
    return NULL;
}

extern "C" HdrImage *sp_FindTexture(sp_MaterialSystem *materialSystem, u32 id)
{
    for (u32 i = 0; i < materialSystem->imageCount; ++i)
        if (materialSystem->imageKeys[i] == id) return materialSystem->images + i;
    return NULL;
}

extern "C" b32 sp_RegisterTexture(sp_MaterialSystem *materialSystem, HdrImage image, u32 id)
{
    if (materialSystem->imageCount >= SP_MAX_IMAGES) return 0;
    u32 index = materialSystem->imageCount++;
    materialSystem->imageKeys[index] = id;
    materialSystem->images[index] = image;
    return 1;
}

static void pack_vertex(const sp_PathVertex &v, float *p)
{
    memcpy(&p[0], &v.materialId, 4);
    p[1] = v.worldPosition.x; p[2] = v.worldPosition.y; p[3] = v.worldPosition.z;
    p[4] = v.outgoingDir.x; p[5] = v.outgoingDir.y; p[6] = v.outgoingDir.z;
    p[7] = v.incomingDir.x; p[8] = v.incomingDir.y; p[9] = v.incomingDir.z;
    p[10] = v.normal.x; p[11] = v.normal.y; p[12] = v.normal.z;
    p[13] = v.uv.x; p[14] = v.uv.y;
}

extern "C" sp_MaterialOutput sp_EvaluateMaterial(sp_MaterialSystem *materialSystem,
                                                 sp_Material *material, sp_PathVertex *vertex)
{
    Library &L = lib();
    std::lock_guard<std::recursive_mutex> lock(L.mutex);
    ensure_init();
    // evaluate `material` itself (it need not be registered): a one-entry table sharing the
    // caller's textures
    sp_MaterialSystem single = *materialSystem;
    single.count = 1;
    single.keys[0] = 0;
    single.materials[0] = *material;
    const DMaterials *dm = upload_materials(&single);
    float packed[15];
    pack_vertex(*vertex, packed);
    L.scratchA.ensure(sizeof(packed));
    L.scratchC.ensure(32);
    SPB_CUDA(cudaMemcpyAsync(L.scratchA.ptr, packed, sizeof(packed), cudaMemcpyHostToDevice, L.stream));
    launch_evaluate_material(kernel_config(), dm, 0, (const float *)L.scratchA.ptr, (float *)L.scratchC.ptr, L.stream);
    SPB_CUDA(cudaGetLastError());
    float out[7];
    SPB_CUDA(cudaMemcpyAsync(out, L.scratchC.ptr, sizeof(out), cudaMemcpyDeviceToHost, L.stream));
    SPB_CUDA(cudaStreamSynchronize(L.stream));
    sp_MaterialOutput r;
    r.albedo = vec3{out[0], out[1], out[2]};
    r.emission = vec3{out[3], out[4], out[5]};
    r.roughness = out[6];
    return r;
}

extern "C" vec3 ComputeRadianceForPath(sp_PathVertex *path, u32 pathLength,
                                       sp_MaterialSystem *materialSystem)
{
    Library &L = lib();
    std::lock_guard<std::recursive_mutex> lock(L.mutex);
    ensure_init();
    const DMaterials *dm = upload_materials(materialSystem);
    std::vector<float> packed((size_t)(pathLength ? pathLength : 1) * 15);
    for (u32 i = 0; i < pathLength; ++i) pack_vertex(path[i], &packed[(size_t)i * 15]);
    L.scratchA.ensure(packed.size() * 4);
    L.scratchC.ensure(16);
    SPB_CUDA(cudaMemcpyAsync(L.scratchA.ptr, packed.data(), packed.size() * 4, cudaMemcpyHostToDevice, L.stream));
    launch_radiance_for_path(kernel_config(), dm, (const float *)L.scratchA.ptr, pathLength,
                             L.params.radianceClamp, (float *)L.scratchC.ptr, L.stream);
    SPB_CUDA(cudaGetLastError());
    float out[3];
    SPB_CUDA(cudaMemcpyAsync(out, L.scratchC.ptr, sizeof(out), cudaMemcpyDeviceToHost, L.stream));
    SPB_CUDA(cudaStreamSynchronize(L.stream));
    return vec3{out[0], out[1], out[2]};
}

// =============================================================================================
// camera, tiles, queue: plain host code with the reference's arithmetic

extern "C" void sp_ConfigureCamera(sp_Camera *camera, ImagePlane *imagePlane, vec3 position,
                                   quat rotation, f32 filmDistance)
{
    configure_camera(camera, imagePlane, position, rotation, filmDistance);
}

extern "C" u32 sp_CalculateFilmPositions(sp_Camera *camera, vec3 *filmPositions,
                                         vec2 *pixelPositions, u32 count)
{
    // simd_path_tracer.cpp:38-63
    ImagePlane *imagePlane = camera->imagePlane;
    for (u32 i = 0; i < count; ++i)
    {
        f32 fx = pixelPositions[i].x / (f32)imagePlane->width;
        f32 fy = pixelPositions[i].y / (f32)imagePlane->height;
        fy = 1.0f - fy;
        fx = fx * 2.0f - 1.0f;
        fy = fy * 2.0f - 1.0f;
        f32 sx = camera->halfFilmWidth * fx, sy = camera->halfFilmHeight * fy;
        vec3 p = {camera->basis.right.x * sx, camera->basis.right.y * sx, camera->basis.right.z * sx};
        p.x = p.x + camera->basis.up.x * sy; p.y = p.y + camera->basis.up.y * sy; p.z = p.z + camera->basis.up.z * sy;
        p.x = p.x + camera->filmCenter.x; p.y = p.y + camera->filmCenter.y; p.z = p.z + camera->filmCenter.z;
        filmPositions[i] = p;
    }
    return count;
}

extern "C" Aabb TransformAabb(vec3 boxMin, vec3 boxMax, vec3 position, quat orientation, vec3 scale)
{
    spbh::V3 lo, hi;
    spbh::transform_aabb(spbh::v3(boxMin.x, boxMin.y, boxMin.z), spbh::v3(boxMax.x, boxMax.y, boxMax.z),
                         spbh::v3(position.x, position.y, position.z),
                         spbh::V4{orientation.x, orientation.y, orientation.z, orientation.w},
                         spbh::v3(scale.x, scale.y, scale.z), &lo, &hi);
    Aabb r;
    r.min = vec3{lo.x, lo.y, lo.z};
    r.max = vec3{hi.x, hi.y, hi.z};
    return r;
}

extern "C" u32 ComputeTiles(u32 totalWidth, u32 totalHeight, u32 tileWidth, u32 tileHeight,
                            Tile *tiles, u32 maxTiles)
{
    // tile.h:11-42 (the counts come from a float Ceil there, hence the float division)
    u32 tileCountY = (u32)ceilf((f32)totalHeight / (f32)tileHeight);
    u32 tileCountX = (u32)ceilf((f32)totalWidth / (f32)tileWidth);
    for (u32 tileY = 0; tileY < tileCountY; ++tileY)
    {
        for (u32 tileX = 0; tileX < tileCountX; ++tileX)
        {
            u32 index = tileX + tileY * tileCountX;
            if (index >= maxTiles) break;
            Tile tile;
            tile.minX = tileX * tileWidth;
            tile.minY = tileY * tileHeight;
            tile.maxX = tile.minX + tileWidth < totalWidth ? tile.minX + tileWidth : totalWidth;
            tile.maxY = tile.minY + tileHeight < totalHeight ? tile.minY + tileHeight : totalHeight;
            tiles[index] = tile;
        }
    }
    u32 total = tileCountY * tileCountX;
    return total < maxTiles ? total : maxTiles;
}

extern "C" WorkQueue CreateWorkQueue(MemoryArena *arena, u32 objectSize, u32 maxObjects)
{
    // work_queue.h:13-22; AllocateBytes from the caller's arena (platform.h:107-114)
    WorkQueue result;
    memset(&result, 0, sizeof(result));
    u64 length = (u64)objectSize * maxObjects;
    SPB_ASSERT(arena->size + length <= arena->capacity);
    result.buffer = (u8 *)arena->base + arena->size;
    arena->size += length;
    result.objectSize = objectSize;
    result.maxObjects = maxObjects;
    return result;
}

extern "C" b32 WorkQueuePush(WorkQueue *queue, void *object, u32 objectSize)
{
    SPB_ASSERT(objectSize == queue->objectSize);
    SPB_ASSERT(queue->tail < (i32)queue->maxObjects);
    memcpy((u8 *)queue->buffer + (size_t)objectSize * queue->tail, object, objectSize);
    queue->tail++;
    return 1;
}

extern "C" void *WorkQueuePop(WorkQueue *queue, u32 objectSize)
{
    SPB_ASSERT(objectSize == queue->objectSize);
    SPB_ASSERT(queue->head != queue->tail);
    i32 index = __atomic_fetch_add(&queue->head, 1, __ATOMIC_SEQ_CST); // intrinsics.h:5-16
    return (u8 *)queue->buffer + (size_t)index * objectSize;
}

// =============================================================================================
// rendering

// What sp_b200_RenderRowsBegin leaves for sp_b200_RenderRowsEnd (one per FrameState).
struct PendingFrame
{
    RenderArgs args;
    DCamera cam;
    u32 rowBegin = 0, rowEnd = 0, tileRows = 0, spp = 0;
    bool wavefront = false, rowsStreamed = false, wantCost = false, pinnedCounters = false;
    uint64_t instancedTriangles = 0;
    f32 *hostPixels = nullptr;
    v4f *image = nullptr;
    unsigned long long *ctr = nullptr; // device counters of the frame
    sp_Scene *sceneHandle = nullptr;
    size_t cCount = 0, waveCount = 0;
    std::vector<unsigned long long> c; // (only when the counters do not fit the pinned landing zone)
    std::vector<uint32_t> waveCounters;
};
static PendingFrame &pending_of(Library &L, FrameState *F)
{
    std::shared_ptr<PendingFrame> &p = L.pendingFrames[F == &L.frames[1] ? 1 : 0];
    if (!p) p = std::make_shared<PendingFrame>();
    return *p;
}

// Enqueues everything a strip needs -- coverage pass, kernels, read-back of the counters, rows to the
// host -- on the render / copy streams and returns without waiting for the device.
static int render_rows_begin(sp_Context *ctx, u32 rowBegin, u32 rowEnd, u32 frame, f32 *hostPixels, void *devicePixels,
                             bool wantCost, FrameState *F)
{
    Library &L = lib();
    L.f = F;
    SPB_ASSERT(ctx && ctx->camera && ctx->camera->imagePlane);
    PendingFrame &P = pending_of(L, F);
    DCamera cam;
    convert_camera(ctx->camera, &cam);
    if (rowEnd > cam.height) rowEnd = cam.height;
    P.rowBegin = rowBegin;
    P.rowEnd = rowEnd;
    P.cam = cam;
    P.tileRows = 0;
    if (rowBegin >= rowEnd || cam.width == 0) return 0;
    DeviceScene *ds = find_scene(ctx->scene);

    size_t imageBytes = (size_t)cam.width * cam.height * 16;
    v4f *image = (v4f *)devicePixels;
    if (!image)
    {
        F->image.ensure(imageBytes);
        image = (v4f *)F->image.ptr;
    }
    u32 tileH = L.params.tileHeight;
    u32 firstTileRow = rowBegin / tileH;
    u32 tileRows = (rowEnd - 1) / tileH - firstTileRow + 1;

    // (the wavefront kernels run the reference's triangle test only: the watertight option renders
    // with the per-pixel kernel, whose walk carries the padded box tests that option needs)
    const bool wavefront = L.params.renderMode != SP_B200_RENDER_PER_PIXEL &&
                           L.params.triangleTest == SP_B200_TRIANGLE_MOLLER_TRUMBORE;
    SPB_CUDA(cudaEventRecord(F->evStart, L.stream));
    // (the wavefront path waits for texture uploads where its first texture-reading kernel starts)
    const DMaterials *dm = upload_materials(ctx->materialSystem, nullptr, wavefront);
    // two cost classes per tile row (queue kernels, sky kernels): 2 x tileRows slots behind the counters
    unsigned long long *ctr = reset_counters((size_t)tileRows * 2);
    F->skyTimed = false;

    RenderArgs &args = P.args;
    args.scene = launch_scene(ds);
    args.materials = dm;
    args.camera = cam;
    args.x0 = 0;
    args.x1 = cam.width;
    args.y0 = rowBegin;
    args.y1 = rowEnd;
    args.spp = L.params.samplesPerPixel;
    args.bounces = L.params.bounceCount;
    args.frame = frame;
    args.clampValue = L.params.radianceClamp;
    args.out = image;
    args.counters = ctr;
    args.tileRowCost = wantCost ? ctr + CTR_COUNT : nullptr;
    args.tileHeight = tileH;
    // the kernel tiles the rectangle from y0 in 16-row CTAs; keep CTA rows inside one tile row
    // by starting at a multiple of 4 (tileHeight % 4 == 0 is enforced by sp_b200_SetParams)
    SPB_ASSERT(rowBegin % 4 == 0 || !wantCost);

    P.tileRows = tileRows;
    P.spp = L.params.samplesPerPixel;
    P.wavefront = wavefront;
    P.wantCost = wantCost;
    P.instancedTriangles = ds->instancedTriangles;
    P.hostPixels = hostPixels;
    P.image = image;
    P.ctr = ctr;
    P.sceneHandle = ctx->scene;
    P.cCount = CTR_COUNT + (size_t)tileRows * 2;
    P.waveCounters.clear();
    SPB_CUDA(cudaEventRecord(F->evKernel0, L.stream));
    if (!wavefront)
        launch_render(kernel_config(), args, L.stream);
    else
        P.rowsStreamed = render_wavefront(args, ds->instancedTriangles, P.waveCounters, hostPixels);
    if (!wavefront) P.rowsStreamed = false;
    SPB_CUDA(cudaGetLastError());
    SPB_CUDA(cudaEventRecord(F->evKernel1, L.stream));

    // counters -> host: pinned landing zones (an asynchronous copy to pageable memory would make this call
    // wait for the whole frame), vectors only when a frame has more counters than the zones hold
    P.waveCount = P.waveCounters.size();
    P.pinnedCounters = P.cCount <= F->hostCountersCap && P.waveCount <= F->hostWaveCap;
    if (!P.pinnedCounters) P.c.resize(P.cCount);
    SPB_CUDA(cudaMemcpyAsync(P.pinnedCounters ? F->hostCounters : P.c.data(), ctr, P.cCount * 8, cudaMemcpyDeviceToHost, L.stream));
    if (P.waveCount)
        SPB_CUDA(cudaMemcpyAsync(P.pinnedCounters ? F->hostWave : P.waveCounters.data(), L.wCtr.ptr, P.waveCount * 4,
                                 cudaMemcpyDeviceToHost, L.stream));
    if (hostPixels && !P.rowsStreamed)
    {
        size_t offset = (size_t)rowBegin * cam.width;
        SPB_CUDA(cudaMemcpyAsync(hostPixels + offset * 4, image + offset,
                                 (size_t)(rowEnd - rowBegin) * cam.width * 16,
                                 cudaMemcpyDeviceToHost, L.stream));
    }
    // (rows that went out on the copy stream: the render stream does NOT wait for them -- the next frame's
    // kernels may start under the tail of this frame's copies; sp_b200_RenderRowsEnd waits for both)
    SPB_CUDA(cudaEventRecord(F->evEnd, L.stream));
    F->pending = true;
    return 0;
}

// Waits for a frame begun above and turns its counters into metrics, statistics and per-row cost.
static int render_rows_end(FrameState *F, sp_Metrics *metrics, u64 *tileRowCost)
{
    Library &L = lib();
    L.f = F;
    PendingFrame &P = pending_of(L, F);
    F->pending = false;
    if (P.tileRows == 0) return 0; // (an empty strip)
    const bool wavefront = P.wavefront;
    const DCamera &cam = P.cam;
    const u32 rowBegin = P.rowBegin, rowEnd = P.rowEnd, tileRows = P.tileRows;
    SPB_CUDA(cudaEventSynchronize(F->evEnd));
    if (P.rowsStreamed) SPB_CUDA(cudaEventSynchronize(F->evCopyDone));
    if (wavefront && F->coverSpeculated && (F->coveredPinned[0] != F->specCovered || F->coveredPinned[1] != F->specHash))
    {
        // (the scene may have been rebuilt since Begin: the arrays the frame was launched on are freed behind it)
        if (P.sceneHandle) P.args.scene = launch_scene(find_scene(P.sceneHandle));
        // the frame was launched on a stale covered-block count / block list (the scene's arrays were
        // rewritten in place, or something this cache's key does not see): render it again on the fresh one
        L.coverCache.valid = false;
        SPB_CUDA(cudaStreamSynchronize(L.stream)); // (a later frame begun on the same stale numbers finishes first; its End renders it again too)
        SPB_CUDA(cudaMemsetAsync(P.ctr, 0, P.cCount * sizeof(unsigned long long), L.stream));
        SPB_CUDA(cudaEventRecord(F->evStart, L.stream));
        SPB_CUDA(cudaEventRecord(F->evKernel0, L.stream));
        P.waveCounters.clear();
        P.rowsStreamed = render_wavefront(P.args, P.instancedTriangles, P.waveCounters, P.hostPixels, false);
        SPB_CUDA(cudaGetLastError());
        SPB_CUDA(cudaEventRecord(F->evKernel1, L.stream));
        P.waveCount = P.waveCounters.size();
        P.pinnedCounters = false;
        P.c.resize(P.cCount);
        SPB_CUDA(cudaMemcpyAsync(P.c.data(), P.ctr, P.cCount * 8, cudaMemcpyDeviceToHost, L.stream));
        if (P.waveCount)
            SPB_CUDA(cudaMemcpyAsync(P.waveCounters.data(), L.wCtr.ptr, P.waveCount * 4, cudaMemcpyDeviceToHost, L.stream));
        if (P.hostPixels && !P.rowsStreamed)
        {
            size_t offset = (size_t)rowBegin * cam.width;
            SPB_CUDA(cudaMemcpyAsync(P.hostPixels + offset * 4, P.image + offset, (size_t)(rowEnd - rowBegin) * cam.width * 16,
                                     cudaMemcpyDeviceToHost, L.stream));
        }
        SPB_CUDA(cudaEventRecord(F->evEnd, L.stream));
        SPB_CUDA(cudaStreamSynchronize(L.stream));
        if (P.rowsStreamed) SPB_CUDA(cudaEventSynchronize(F->evCopyDone));
    }
    unsigned long long *c = P.pinnedCounters ? F->hostCounters : P.c.data();
    const uint32_t *waveCounters = P.pinnedCounters ? F->hostWave : P.waveCounters.data();
    const size_t waveCount = P.waveCount;
    float kernelMs = 0, totalMs = 0;
    SPB_CUDA(cudaEventElapsedTime(&kernelMs, F->evKernel0, F->evKernel1));
    SPB_CUDA(cudaEventElapsedTime(&totalMs, F->evStart, F->evEnd));
    if (P.rowsStreamed)
    {
        // the copies' tail beyond the last kernel belongs to the call's device time
        float tail = 0.0f;
        if (cudaEventElapsedTime(&tail, F->evEnd, F->evCopyDone) == cudaSuccess && tail > 0.0f) totalMs += tail;
        else cudaGetLastError();
    }
    float traceMs = 0.0f;
    if (wavefront)
        for (size_t i = 0; i + 1 < F->traceEventsUsed; i += 2)
        {
            float ms = 0.0f;
            SPB_CUDA(cudaEventElapsedTime(&ms, F->traceEvents[i], F->traceEvents[i + 1]));
            traceMs += ms;
        }
    if (wavefront && !L.tuner.probes.empty())
    {
        // measurements of the sorted-class refill threshold (render_wavefront): ms per ray of the
        // bounce-1 trace launches that were timed; decide once both candidates have been seen
        Library::SortedTuner &tn = L.tuner;
        for (const Library::SortedTuner::Probe &pr : tn.probes)
        {
            float ms = 0.0f;
            SPB_CUDA(cudaEventElapsedTime(&ms, pr.e0, pr.e1));
            double rays = pr.ctrIndex + WCTR_NHITS < waveCount ? (double)waveCounters[pr.ctrIndex + WCTR_NHITS] : 0.0;
            if (rays < 1.0) continue;
            tn.ms[pr.candidate] += ms;
            tn.rays[pr.candidate] += rays;
            tn.samples[pr.candidate]++;
        }
        tn.probes.clear();
        tn.frames++;
        // one sample each is enough (the candidates differ by ~30 % where it matters), so a strip
        // with a single pass per frame has decided after two warm-up frames
        if (tn.samples[0] >= 1 && tn.samples[1] >= 1)
            tn.choice = tn.ms[0] / tn.rays[0] <= tn.ms[1] / tn.rays[1] ? 0 : 1;
    }
    if (wavefront)
    {
        // queue lengths are the path counters: rays = rays entering each traversal, and every
        // ray ends in exactly one of the hit / miss queues; a sky pixel is spp rays, spp misses
        unsigned long long rays = 0, hits = 0, misses = 0;
        for (size_t i = 0; i + WCTR_STRIDE <= waveCount; i += WCTR_STRIDE)
        {
            hits += waveCounters[i + WCTR_NHITS];
            misses += waveCounters[i + WCTR_NMISSES];
        }
        // a sky-kernel pixel is spp rays, spp misses (its row cost was added on the device)
        misses += c[CTR_SKY_PIXELS] * P.spp;
        rays = hits + misses;
        c[CTR_PATHS] = (unsigned long long)(rowEnd - rowBegin) * cam.width * P.spp;
        c[CTR_RAYS] = rays;
        c[CTR_HITS] = hits;
        c[CTR_MISSES] = misses;
    }
    add_metrics(metrics, c, kernelMs);
    record_stats(c, kernelMs, totalMs);
    L.lastStats.traceMs = wavefront ? traceMs : kernelMs;
    L.lastStats.traceLaunches = wavefront ? (u32)(F->traceEventsUsed / 2) : 1u;
    L.lastStats.tracedRays = c[CTR_RAYS] - (wavefront ? c[CTR_SKY_PIXELS] * P.spp : 0);
    if (tileRowCost && P.wantCost)
    {
        // Cost of every tile row in NANOSECONDS of this device: the units the kernels counted,
        // converted class by class with the time that class took in this very call -- the sky
        // kernels' units by the sky kernels' time, the queue kernels' (escaped rays, surface hits)
        // by the rest.  A unit of sky is not a unit of bunny, and how much not depends on the
        // scene and the strip; measuring the two apart takes that guess out of the re-cut.
        float skyMs = 0.0f;
        if (wavefront && F->skyTimed) SPB_CUDA(cudaEventElapsedTime(&skyMs, F->evSky0, F->evSky1));
        double unitsQueue = 0.0, unitsSky = 0.0;
        for (u32 i = 0; i < tileRows; ++i)
        {
            unitsQueue += (double)c[CTR_COUNT + i];
            unitsSky += (double)c[CTR_COUNT + tileRows + i];
        }
        double nsSky = (double)skyMs * 1e6, nsQueue = (double)kernelMs * 1e6 - nsSky;
        if (nsQueue < 0.0) nsQueue = 0.0;
        if (unitsSky <= 0.0) { nsQueue += nsSky; nsSky = 0.0; }
        if (unitsQueue <= 0.0) { nsSky += nsQueue; nsQueue = 0.0; }
        const double perQueue = unitsQueue > 0.0 ? nsQueue / unitsQueue : 0.0, perSky = unitsSky > 0.0 ? nsSky / unitsSky : 0.0;
        for (u32 i = 0; i < tileRows; ++i)
            tileRowCost[i] = (u64)((double)c[CTR_COUNT + i] * perQueue + (double)c[CTR_COUNT + tileRows + i] * perSky + 0.5);
    }
    return 0;
}

extern "C" int sp_b200_RenderRows(sp_Context *ctx, u32 rowBegin, u32 rowEnd, u32 frame,
                                  f32 *hostPixels, void *devicePixels, sp_Metrics *metrics,
                                  u64 *tileRowCost)
{
    Library &L = lib();
    std::lock_guard<std::recursive_mutex> lock(L.mutex);
    ensure_init();
    FrameState *F = !L.frames[0].pending ? &L.frames[0] : (!L.frames[1].pending ? &L.frames[1] : nullptr);
    if (!F)
    {
        log_message("sp_b200_RenderRows: two frames are already in flight (sp_b200_RenderRowsBegin): end one first");
        return -1;
    }
    if (render_rows_begin(ctx, rowBegin, rowEnd, frame, hostPixels, devicePixels, tileRowCost != nullptr, F) != 0) return -1;
    return render_rows_end(F, metrics, tileRowCost);
}

extern "C" int sp_b200_RenderRowsBegin(sp_Context *ctx, u32 rowBegin, u32 rowEnd, u32 frame,
                                       f32 *hostPixels, void *devicePixels, int wantTileRowCost)
{
    Library &L = lib();
    std::lock_guard<std::recursive_mutex> lock(L.mutex);
    ensure_init();
    int slot = !L.frames[0].pending ? 0 : (!L.frames[1].pending ? 1 : -1);
    if (slot < 0) return -1;
    // the frame begun before this one, if any, is the other slot: alternate, so that frames end in the order they began
    if (render_rows_begin(ctx, rowBegin, rowEnd, frame, hostPixels, devicePixels, wantTileRowCost != 0, &L.frames[slot]) != 0) return -1;
    L.frames[slot].pending = true; // (also for an empty strip: End is due either way)
    return slot;
}

extern "C" int sp_b200_RenderRowsEnd(int slot, sp_Metrics *metrics, u64 *tileRowCost)
{
    Library &L = lib();
    std::lock_guard<std::recursive_mutex> lock(L.mutex);
    if (slot < 0 || slot > 1 || !L.frames[slot].pending) return -1;
    SPB_CUDA(cudaSetDevice(L.device)); // (the calling thread need not be the one that began the frame)
    return render_rows_end(&L.frames[slot], metrics, tileRowCost);
}


// Output stage: tone mapping + 8-bit store (post_processing.frag.glsl:19-26).  Source: device
// pixels if given, else host pixels (uploaded); result to hostRGBA8 and/or deviceRGBA8.
extern "C" int sp_b200_ToneMap(const f32 *hostPixels, const void *devicePixels, u32 pixelCount, f32 exposure,
                               u32 *hostRGBA8, void *deviceRGBA8)
{
    Library &L = lib();
    std::lock_guard<std::recursive_mutex> lock(L.mutex);
    ensure_init();
    if (pixelCount == 0) return 0;
    SPB_ASSERT(hostPixels || devicePixels);
    SPB_ASSERT(hostRGBA8 || deviceRGBA8);
    const v4f *src = (const v4f *)devicePixels;
    if (!src)
    {
        L.scratchA.ensure((size_t)pixelCount * 16);
        SPB_CUDA(cudaMemcpyAsync(L.scratchA.ptr, hostPixels, (size_t)pixelCount * 16, cudaMemcpyHostToDevice, L.stream));
        src = (const v4f *)L.scratchA.ptr;
    }
    uint32_t *dst = (uint32_t *)deviceRGBA8;
    if (!dst)
    {
        L.scratchB.ensure((size_t)pixelCount * 4);
        dst = (uint32_t *)L.scratchB.ptr;
    }
    launch_tone_map(kernel_config(), src, pixelCount, exposure, dst, L.stream);
    SPB_CUDA(cudaGetLastError());
    if (hostRGBA8)
        SPB_CUDA(cudaMemcpyAsync(hostRGBA8, dst, (size_t)pixelCount * 4, cudaMemcpyDeviceToHost, L.stream));
    SPB_CUDA(cudaStreamSynchronize(L.stream));
    return 0;
}

// Environment pre-processing (src/cubemap.cpp): the equirectangular map the path samples, turned
// into the cube map / irradiance cube map the reference bakes at start-up (main.cpp:1307-1315).
// The source image goes through the same device texture cache as the path tracer's textures.
static DImage device_image(const HdrImage *equirect)
{
    SPB_ASSERT(equirect && equirect->pixels && equirect->width > 0 && equirect->height > 0);
    DImage env;
    env.pixels = device_texture(*equirect);
    wait_textures();
    env.width = equirect->width;
    env.height = equirect->height;
    env.pad = 0;
    return env;
}

static void finish_faces(const v4f *faces, size_t texels, f32 *hostFaces, void *deviceFaces)
{
    Library &L = lib();
    if (hostFaces)
        SPB_CUDA(cudaMemcpyAsync(hostFaces, faces, texels * 16, cudaMemcpyDeviceToHost, L.stream));
    (void)deviceFaces;
    SPB_CUDA(cudaStreamSynchronize(L.stream));
}

extern "C" int sp_b200_CreateCubeMap(const HdrImage *equirect, u32 width, u32 height, f32 *hostFaces,
                                     void *deviceFaces)
{
    Library &L = lib();
    std::lock_guard<std::recursive_mutex> lock(L.mutex);
    ensure_init();
    SPB_ASSERT(hostFaces || deviceFaces);
    size_t texels = (size_t)6 * width * height;
    if (texels == 0) return 0;
    DImage env = device_image(equirect);
    v4f *faces = (v4f *)deviceFaces;
    if (!faces)
    {
        L.scratchB.ensure(texels * 16);
        faces = (v4f *)L.scratchB.ptr;
    }
    launch_cube_map(kernel_config(), env, faces, width, height, L.stream);
    SPB_CUDA(cudaGetLastError());
    finish_faces(faces, texels, hostFaces, deviceFaces);
    return 0;
}

extern "C" int sp_b200_CreateIrradianceCubeMap(const HdrImage *equirect, u32 width, u32 height,
                                               u32 samplesPerPixel, u32 sampling, f32 sampleDelta,
                                               f32 *hostFaces, void *deviceFaces)
{
    Library &L = lib();
    std::lock_guard<std::recursive_mutex> lock(L.mutex);
    ensure_init();
    SPB_ASSERT(hostFaces || deviceFaces);
    SPB_ASSERT(sampling == SP_B200_IRRADIANCE_UNIFORM || sampling == SP_B200_IRRADIANCE_RANDOM);
    size_t texels = (size_t)6 * width * height;
    if (texels == 0) return 0;
    SPB_ASSERT(texels < 0xFFFFFFFFull);
    IrradianceArgs a;
    memset(&a, 0, sizeof(a));
    a.env = device_image(equirect);
    a.width = width;
    a.height = height;
    a.clampValue = (f32)10; // RADIANCE_CLAMP, config.h:36 (compile-time in the reference's baker)
    std::vector<uint32_t> staging;
    if (sampling == SP_B200_IRRADIANCE_UNIFORM)
    {
        // the reference's loop variables (cubemap.cpp:162-164): phi and theta grow by repeated
        // float addition of sampleDelta (0.1f there), bounds 2*PI and 0.5*PI in float
        SPB_ASSERT(sampleDelta > 1e-3f);
        std::vector<float> phis(irradiance_loop_values(2.0f * SPB_PI, sampleDelta, nullptr, 0));
        std::vector<float> thetas(irradiance_loop_values(0.5f * SPB_PI, sampleDelta, nullptr, 0));
        irradiance_loop_values(2.0f * SPB_PI, sampleDelta, phis.data(), (uint32_t)phis.size());
        irradiance_loop_values(0.5f * SPB_PI, sampleDelta, thetas.data(), (uint32_t)thetas.size());
        a.phiCount = (uint32_t)phis.size();
        a.thetaCount = (uint32_t)thetas.size();
        staging.resize(phis.size() + thetas.size());
        memcpy(staging.data(), phis.data(), phis.size() * 4);
        memcpy(staging.data() + phis.size(), thetas.data(), thetas.size() * 4);
    }
    else
    {
        SPB_ASSERT(samplesPerPixel > 0 && samplesPerPixel < (1u << 28));
        a.samplesPerPixel = samplesPerPixel;
        a.sampleContribution = 1.0f / (f32)samplesPerPixel; // cubemap.cpp:115
        a.seed = 0x45BA12F3u;                               // cubemap.cpp:123
        staging.resize(2 * 32 * 32);
        xorshift_build_jump_table(3 * samplesPerPixel, staging.data());
        xorshift_build_jump_table(3, staging.data() + 32 * 32);
    }
    L.scratchA.ensure(staging.size() * 4);
    SPB_CUDA(cudaMemcpyAsync(L.scratchA.ptr, staging.data(), staging.size() * 4, cudaMemcpyHostToDevice, L.stream));
    if (sampling == SP_B200_IRRADIANCE_UNIFORM)
    {
        a.phis = (const float *)L.scratchA.ptr;
        a.thetas = a.phis + a.phiCount;
    }
    else
    {
        a.jumpTexel = (const uint32_t *)L.scratchA.ptr;
        a.jumpSample = a.jumpTexel + 32 * 32;
    }
    a.out = (v4f *)deviceFaces;
    if (!a.out)
    {
        L.scratchB.ensure(texels * 16);
        a.out = (v4f *)L.scratchB.ptr;
    }
    launch_irradiance(kernel_config(), a, sampling == SP_B200_IRRADIANCE_RANDOM ? 1 : 0, L.stream);
    SPB_CUDA(cudaGetLastError());
    finish_faces(a.out, texels, hostFaces, deviceFaces);
    return 0;
}

// One frame over every device of sp_b200_InitDevices.  Each device renders a strip of rows cut at
// multiples of tileHeight: the primary on the calling thread, the others on their worker threads,
// all through sp_b200_RenderRows.  Output: hostPixels (each device copies its rows there itself,
// over its own PCIe link) and / or primaryDevicePixels (the other devices send their strips with
// peer copies: the gather of SURVEY.md §8e).  The cut of the NEXT frame is made from what this one
// cost: per-row cost units scaled by each device's kernel time (spb_strips.h).
static int render_frame_devices(sp_Context *ctx, u32 frame, f32 *hostPixels, void *primaryDevicePixels, sp_Metrics *metrics)
{
    MultiDevice &M = multi();
    std::lock_guard<std::mutex> lock(M.mutex);
    Library &P = primary_lib();
    const u32 parts = (u32)M.devices.size();
    const u32 height = ctx->camera->imagePlane->height, width = ctx->camera->imagePlane->width;
    const u32 quantum = P.params.tileHeight;
    const u32 rows = (height + quantum - 1) / quantum;
    if (M.nextBounds.size() == parts + 1 && M.height == height && M.quantum == quantum) M.bounds = M.nextBounds;
    else
    {
        M.bounds.assign(parts + 1, 0);
        partition_rows(height, quantum, parts, nullptr, M.bounds.data());
        M.height = height;
        M.quantum = quantum;
    }
    M.stats.assign(parts, sp_b200_Stats());
    M.seconds.assign(parts, 0.0);
    std::vector<sp_Metrics> partMetrics(parts);
    std::vector<std::vector<u64>> partCost(parts);
    std::vector<double> units(rows, 0.0);
    memset(partMetrics.data(), 0, sizeof(sp_Metrics) * parts);
    // what the host set on the primary holds for every device
    struct Settings
    {
        sp_b200_Params params; bool stats, sky, sort, cand, one, overlap; int fuse; uint32_t paths, sortBounces, refill[3], evict[2];
    } set = {P.params, P.statsEnabled, P.skyCulling, P.sortBounceRays, P.primaryCandidates, P.skyOneLookup, P.overlapCopies, P.fuseMiss,
             P.pathsPerPass, P.sortBounces, {P.refillThreshold[0], P.refillThreshold[1], P.refillThreshold[2]}, {P.evictBelow[0], P.evictBelow[1]}};
    const int primaryDevice = M.devices[0];
    auto render_part = [&](u32 p) {
        const u32 b = M.bounds[p], e = M.bounds[p + 1];
        if (e <= b) return;
        Library &L = lib();
        if (p > 0)
        {
            L.params = set.params; L.statsEnabled = set.stats; L.skyCulling = set.sky; L.sortBounceRays = set.sort;
            L.primaryCandidates = set.cand; L.skyOneLookup = set.one; L.overlapCopies = set.overlap; L.fuseMiss = set.fuse;
            L.pathsPerPass = set.paths; L.sortBounces = set.sortBounces;
            for (int k = 0; k < 3; ++k) L.refillThreshold[k] = set.refill[k];
            L.evictBelow[0] = set.evict[0]; L.evictBelow[1] = set.evict[1];
        }
        partCost[p].assign((e - 1) / quantum - b / quantum + 1, 0);
        // the primary renders straight into the gathered image; the others into their own
        void *device = p == 0 ? primaryDevicePixels : nullptr;
        sp_b200_RenderRows(ctx, b, e, frame, hostPixels, device, &partMetrics[p], partCost[p].data());
        if (p > 0 && primaryDevicePixels)
        {
            const size_t offset = (size_t)b * width * 16, bytes = (size_t)(e - b) * width * 16;
            SPB_CUDA(cudaMemcpyPeerAsync((char *)primaryDevicePixels + offset, primaryDevice, (const char *)L.f->image.ptr + offset,
                                         L.device, bytes, L.stream));
            SPB_CUDA(cudaStreamSynchronize(L.stream));
        }
        M.stats[p] = L.lastStats;
        M.seconds[p] = (double)L.lastStats.kernelMs * 1e-3;
    };
    for (u32 p = 1; p < parts; ++p) M.workers[p - 1]->post([&render_part, p] { render_part(p); });
    render_part(0);
    for (u32 p = 1; p < parts; ++p) M.workers[p - 1]->wait();

    // metrics: counters add up, "cycles" (device nanoseconds) are the slowest device's
    u64 slowest = 0;
    for (u32 p = 0; p < parts; ++p)
    {
        if (metrics)
            for (int k = 0; k < 12; ++k)
                if (k != sp_Metric_CyclesElapsed) metrics->values[k] += partMetrics[p].values[k];
        if (partMetrics[p].values[sp_Metric_CyclesElapsed] > slowest) slowest = partMetrics[p].values[sp_Metric_CyclesElapsed];
        const u32 b = M.bounds[p], e = M.bounds[p + 1];
        if (e <= b) continue;
        const u32 first = b / quantum;
        for (size_t i = 0; i < partCost[p].size() && first + i < rows; ++i) units[first + i] += (double)partCost[p][i];
    }
    if (metrics) metrics->values[sp_Metric_CyclesElapsed] = slowest;
    // the whole frame's figures on the primary (sp_b200_GetLastStats); per device: sp_b200_GetDeviceStats
    sp_b200_Stats total = M.stats[0];
    for (u32 p = 1; p < parts; ++p)
    {
        const sp_b200_Stats &s = M.stats[p];
        total.rays += s.rays; total.nodeVisits += s.nodeVisits; total.triangleTests += s.triangleTests;
        total.objectTests += s.objectTests; total.envClampedLookups += s.envClampedLookups;
        total.tracedRays += s.tracedRays; total.traceLaunches += s.traceLaunches;
        if (s.kernelMs > total.kernelMs) total.kernelMs = s.kernelMs;
        if (s.totalMs > total.totalMs) total.totalMs = s.totalMs;
        if (s.traceMs > total.traceMs) total.traceMs = s.traceMs;
    }
    P.lastStats = total;
    // next frame's cut
    std::vector<double> cost = row_seconds(height, quantum, parts, M.bounds.data(), units.data(), M.seconds.data());
    double sum = 0.0;
    for (double c : cost) sum += c;
    // (M.bounds / M.stats keep describing the frame just rendered: sp_b200_GetDeviceStats)
    M.nextBounds = M.bounds;
    if (sum > 0.0 && rows > parts) partition_rows(height, quantum, parts, cost.data(), M.nextBounds.data());
    return 0;
}

extern "C" int sp_b200_RenderFrame(sp_Context *ctx, u32 frame, sp_Metrics *metrics)
{
    SPB_ASSERT(ctx && ctx->camera && ctx->camera->imagePlane);
    ImagePlane *plane = ctx->camera->imagePlane;
    if (multi().active() && !t_current) return render_frame_devices(ctx, frame, (f32 *)plane->pixels, nullptr, metrics);
    return sp_b200_RenderRows(ctx, 0, plane->height, frame, (f32 *)plane->pixels, nullptr, metrics, nullptr);
}

extern "C" int sp_b200_RenderFrameToDevice(sp_Context *ctx, u32 frame, void *devicePixels, sp_Metrics *metrics)
{
    SPB_ASSERT(ctx && ctx->camera && ctx->camera->imagePlane && devicePixels);
    ImagePlane *plane = ctx->camera->imagePlane;
    if (multi().active() && !t_current) return render_frame_devices(ctx, frame, nullptr, devicePixels, metrics);
    return sp_b200_RenderRows(ctx, 0, plane->height, frame, nullptr, devicePixels, metrics, nullptr);
}

// sp_PathTraceTile is re-entrant in the reference: WorkerThread (main.cpp:728-759) calls it from 16
// host threads at once, one tile each.  A tile is one GPU thread of control by definition (its pixels
// share one serial XorShift32 stream), so a launch per call would keep one lane of the whole GPU busy
// and 15 host threads waiting on the library mutex.  Concurrent callers are COMBINED instead: a caller
// that finds nobody rendering becomes the leader and renders every request that has queued up behind
// it -- its own first, then, while that launch runs, the requests of the other threads, all of one
// context in ONE launch of k_tiles_serial -- until the queue is empty; the others sleep until their
// request is marked done.  Results are per tile (own rng state, own pixels, own metrics), so a caller
// cannot tell whether it was batched.
struct TileRequest
{
    sp_Context *ctx;
    Tile tile;
    RandomNumberGenerator *rng;
    sp_Metrics *metrics;
    bool done;
};

struct TileCombiner
{
    std::mutex m;
    std::condition_variable cv;
    std::vector<TileRequest *> pending;
    bool leaderActive = false;
    unsigned long long launches = 0, tiles = 0;
};

static TileCombiner &tile_combiner()
{
    static TileCombiner *instance = new TileCombiner();
    return *instance;
}

// every request of `group` has the same context: one launch
static void render_tile_group(const std::vector<TileRequest *> &group)
{
    Library &L = lib();
    std::lock_guard<std::recursive_mutex> lock(L.mutex);
    ensure_init();
    sp_Context *ctx = group[0]->ctx;
    SPB_ASSERT(ctx && ctx->camera && ctx->camera->imagePlane);
    ImagePlane *plane = ctx->camera->imagePlane;
    const u32 count = (u32)group.size();
    std::vector<uint32_t> staging((size_t)count * 5); // per tile: minX minY maxX maxY, then one rng state per tile
    for (u32 k = 0; k < count; ++k)
    {
        const Tile &t = group[k]->tile;
        staging[(size_t)k * 4 + 0] = t.minX;
        staging[(size_t)k * 4 + 1] = t.minY;
        staging[(size_t)k * 4 + 2] = t.maxX < plane->width ? t.maxX : plane->width; // simd_path_tracer.cpp:188-191
        staging[(size_t)k * 4 + 3] = t.maxY < plane->height ? t.maxY : plane->height;
        staging[(size_t)count * 4 + k] = group[k]->rng->state;
    }
    DCamera cam;
    convert_camera(ctx->camera, &cam);
    DeviceScene *ds = find_scene(ctx->scene);
    size_t imageBytes = (size_t)cam.width * cam.height * 16;
    L.f->image.ensure(imageBytes ? imageBytes : 16);
    SPB_CUDA(cudaEventRecord(L.f->evStart, L.stream));
    const DMaterials *dm = upload_materials(ctx->materialSystem);
    unsigned long long *ctr = reset_counters((size_t)(count - 1) * CTR_COUNT);
    L.scratchA.ensure(staging.size() * 4);
    SPB_CUDA(cudaMemcpyAsync(L.scratchA.ptr, staging.data(), staging.size() * 4, cudaMemcpyHostToDevice, L.stream));

    TileArgs args;
    args.scene = launch_scene(ds);
    args.materials = dm;
    args.camera = cam;
    args.tiles = (const uint32_t *)L.scratchA.ptr;
    args.rngStates = (uint32_t *)L.scratchA.ptr + (size_t)count * 4;
    args.count = count;
    args.spp = L.params.samplesPerPixel;
    args.bounces = L.params.bounceCount;
    args.clampValue = L.params.radianceClamp;
    args.out = (v4f *)L.f->image.ptr;
    args.counters = ctr;
    SPB_CUDA(cudaEventRecord(L.f->evKernel0, L.stream));
    launch_tiles_serial(kernel_config(), args, L.stream);
    SPB_CUDA(cudaGetLastError());
    SPB_CUDA(cudaEventRecord(L.f->evKernel1, L.stream));

    std::vector<unsigned long long> c((size_t)count * CTR_COUNT);
    std::vector<uint32_t> states(count);
    SPB_CUDA(cudaMemcpyAsync(c.data(), ctr, c.size() * 8, cudaMemcpyDeviceToHost, L.stream));
    SPB_CUDA(cudaMemcpyAsync(states.data(), (uint32_t *)L.scratchA.ptr + (size_t)count * 4, (size_t)count * 4,
                             cudaMemcpyDeviceToHost, L.stream));
    if (plane->pixels)
        for (u32 k = 0; k < count; ++k)
        {
            const uint32_t *t = staging.data() + (size_t)k * 4;
            if (t[2] <= t[0] || t[3] <= t[1]) continue;
            size_t pitch = (size_t)cam.width * 16;
            size_t offset = (size_t)t[1] * cam.width + t[0];
            SPB_CUDA(cudaMemcpy2DAsync((f32 *)plane->pixels + offset * 4, pitch, (v4f *)L.f->image.ptr + offset,
                                       pitch, (size_t)(t[2] - t[0]) * 16, t[3] - t[1],
                                       cudaMemcpyDeviceToHost, L.stream));
        }
    SPB_CUDA(cudaEventRecord(L.f->evEnd, L.stream));
    SPB_CUDA(cudaStreamSynchronize(L.stream));
    float kernelMs = 0, totalMs = 0;
    SPB_CUDA(cudaEventElapsedTime(&kernelMs, L.f->evKernel0, L.f->evKernel1));
    SPB_CUDA(cudaEventElapsedTime(&totalMs, L.f->evStart, L.f->evEnd));
    int clockKHz = 0;
    cudaDeviceGetAttribute(&clockKHz, cudaDevAttrClockRate, L.device);
    std::vector<unsigned long long> total(CTR_COUNT, 0);
    for (u32 k = 0; k < count; ++k)
    {
        const unsigned long long *ck = c.data() + (size_t)k * CTR_COUNT;
        for (int j = 0; j < CTR_COUNT; ++j) total[j] += ck[j];
        group[k]->rng->state = states[k];
        // CyclesElapsed = this tile's own device time (nanoseconds) when it shared the launch
        float tileMs = (count > 1 && clockKHz > 0) ? (float)((double)ck[CTR_CLOCK_SUM] / (double)clockKHz) : kernelMs;
        add_metrics(group[k]->metrics, ck, tileMs);
    }
    record_stats(total.data(), kernelMs, totalMs);
}

extern "C" void sp_PathTraceTile(sp_Context *ctx, Tile tile, RandomNumberGenerator *rng,
                                 sp_Metrics *metrics)
{
    SPB_ASSERT(ctx && ctx->camera && ctx->camera->imagePlane && rng);
    TileCombiner &C = tile_combiner();
    TileRequest req = {ctx, tile, rng, metrics, false};
    std::unique_lock<std::mutex> lock(C.m);
    C.pending.push_back(&req);
    if (C.leaderActive)
    {
        C.cv.wait(lock, [&] { return req.done; });
        return;
    }
    C.leaderActive = true;
    while (!C.pending.empty())
    {
        std::vector<TileRequest *> batch;
        batch.swap(C.pending);
        lock.unlock();
        // one launch per context, contexts in first-seen order
        std::vector<bool> taken(batch.size(), false);
        for (size_t first = 0; first < batch.size(); ++first)
        {
            if (taken[first]) continue;
            std::vector<TileRequest *> group;
            for (size_t i = first; i < batch.size(); ++i)
                if (!taken[i] && batch[i]->ctx == batch[first]->ctx)
                {
                    taken[i] = true;
                    group.push_back(batch[i]);
                }
            render_tile_group(group);
        }
        lock.lock();
        C.launches++;
        C.tiles += batch.size();
        for (TileRequest *r : batch) r->done = true;
        C.cv.notify_all();
    }
    C.leaderActive = false;
}

extern "C" void sp_b200_TileCombinerStats(u64 *launches, u64 *tiles)
{
    TileCombiner &C = tile_combiner();
    std::lock_guard<std::mutex> lock(C.m);
    if (launches) *launches = C.launches;
    if (tiles) *tiles = C.tiles;
}

// ---------------------------------------------------------------------------------------------
// The reference's tile scheduler (main.cpp:246-250 sp_Task, :728-759 WorkerThread + g_metricsBuffer,
// :819-844 AddRayTracingWorkQueue) on top of the WorkQueue above.  The reference's 16 worker
// threads each pop a task, reseed (0xF51C0E49) and render its tile; here the drain pops every task
// and renders all tiles of one context in ONE launch of the serial-stream tile kernel (one thread
// of control per tile, as the per-tile RNG stream demands), so the pixels are the reference's.

extern "C" u32 sp_b200_AddRayTracingWorkQueue(WorkQueue *workQueue, sp_Context *ctx)
{
    Library &L = lib();
    SPB_ASSERT(workQueue && ctx && ctx->camera && ctx->camera->imagePlane);
    SPB_ASSERT(workQueue->head == workQueue->tail);      // main.cpp:821
    SPB_ASSERT(workQueue->objectSize == sizeof(sp_Task));
    ImagePlane *plane = ctx->camera->imagePlane;
    u32 tileWidth, tileHeight;
    {
        std::lock_guard<std::recursive_mutex> lock(L.mutex);
        tileWidth = L.params.tileWidth ? L.params.tileWidth : 64;    // TILE_WIDTH, config.h:15-16
        tileHeight = L.params.tileHeight ? L.params.tileHeight : 64;
    }
    // MAX_TILES (config.h:19) bounds the reference's stack array; the queue's capacity bounds ours
    std::vector<Tile> tiles(workQueue->maxObjects);
    u32 tileCount = ComputeTiles(plane->width, plane->height, tileWidth, tileHeight, tiles.data(),
                                 workQueue->maxObjects);
    workQueue->tail = 0; // main.cpp:833
    for (u32 i = 0; i < tileCount; ++i)
    {
        sp_Task task;
        memset(&task, 0, sizeof(task));
        task.context = ctx;
        task.tile = tiles[i];
        WorkQueuePush(workQueue, &task, sizeof(task));
    }
    workQueue->head = 0; // main.cpp:843
    return tileCount;
}

extern "C" u32 sp_b200_DrainRayTracingWorkQueue(WorkQueue *queue, sp_Metrics *metricsBuffer, u32 maxMetrics)
{
    Library &L = lib();
    std::lock_guard<std::recursive_mutex> lock(L.mutex);
    ensure_init();
    SPB_ASSERT(queue && queue->objectSize == sizeof(sp_Task));
    std::vector<sp_Task> tasks;
    while (queue->head != queue->tail) // WorkerThread's test, main.cpp:736
        tasks.push_back(*(sp_Task *)WorkQueuePop(queue, sizeof(sp_Task)));
    u32 done = 0;
    // tasks of one context are rendered together; contexts in first-seen order
    std::vector<bool> taken(tasks.size(), false);
    for (size_t first = 0; first < tasks.size(); ++first)
    {
        if (taken[first]) continue;
        sp_Context *ctx = tasks[first].context;
        SPB_ASSERT(ctx && ctx->camera && ctx->camera->imagePlane);
        ImagePlane *plane = ctx->camera->imagePlane;
        std::vector<uint32_t> staging; // per tile: minX minY maxX maxY, then one rng state per tile
        std::vector<size_t> members;
        for (size_t i = first; i < tasks.size(); ++i)
            if (!taken[i] && tasks[i].context == ctx)
            {
                taken[i] = true;
                members.push_back(i);
                const Tile &t = tasks[i].tile;
                staging.push_back(t.minX);
                staging.push_back(t.minY);
                staging.push_back(t.maxX < plane->width ? t.maxX : plane->width);   // simd_path_tracer.cpp:188-191
                staging.push_back(t.maxY < plane->height ? t.maxY : plane->height);
            }
        const u32 count = (u32)members.size();
        staging.resize((size_t)count * 5, 0xF51C0E49u); // rng.state per task, main.cpp:738-739

        DCamera cam;
        convert_camera(ctx->camera, &cam);
        DeviceScene *ds = find_scene(ctx->scene);
        size_t imageBytes = (size_t)cam.width * cam.height * 16;
        L.f->image.ensure(imageBytes ? imageBytes : 16);
        SPB_CUDA(cudaEventRecord(L.f->evStart, L.stream));
        const DMaterials *dm = upload_materials(ctx->materialSystem);
        unsigned long long *ctr = reset_counters((size_t)(count - 1) * CTR_COUNT);
        L.scratchA.ensure(staging.size() * 4);
        SPB_CUDA(cudaMemcpyAsync(L.scratchA.ptr, staging.data(), staging.size() * 4, cudaMemcpyHostToDevice, L.stream));

        TileArgs args;
        args.scene = launch_scene(ds);
        args.materials = dm;
        args.camera = cam;
        args.tiles = (const uint32_t *)L.scratchA.ptr;
        args.rngStates = (uint32_t *)L.scratchA.ptr + (size_t)count * 4;
        args.count = count;
        args.spp = L.params.samplesPerPixel;
        args.bounces = L.params.bounceCount;
        args.clampValue = L.params.radianceClamp;
        args.out = (v4f *)L.f->image.ptr;
        args.counters = ctr;
        SPB_CUDA(cudaEventRecord(L.f->evKernel0, L.stream));
        launch_tiles_serial(kernel_config(), args, L.stream);
        SPB_CUDA(cudaGetLastError());
        SPB_CUDA(cudaEventRecord(L.f->evKernel1, L.stream));

        std::vector<unsigned long long> c((size_t)count * CTR_COUNT);
        SPB_CUDA(cudaMemcpyAsync(c.data(), ctr, c.size() * 8, cudaMemcpyDeviceToHost, L.stream));
        if (plane->pixels)
            for (u32 k = 0; k < count; ++k)
            {
                const uint32_t *t = staging.data() + (size_t)k * 4;
                if (t[2] <= t[0] || t[3] <= t[1]) continue;
                size_t pitch = (size_t)cam.width * 16;
                size_t offset = (size_t)t[1] * cam.width + t[0];
                SPB_CUDA(cudaMemcpy2DAsync((f32 *)plane->pixels + offset * 4, pitch, (v4f *)L.f->image.ptr + offset,
                                           pitch, (size_t)(t[2] - t[0]) * 16, t[3] - t[1],
                                           cudaMemcpyDeviceToHost, L.stream));
            }
        SPB_CUDA(cudaEventRecord(L.f->evEnd, L.stream));
        SPB_CUDA(cudaStreamSynchronize(L.stream));
        float kernelMs = 0, totalMs = 0;
        SPB_CUDA(cudaEventElapsedTime(&kernelMs, L.f->evKernel0, L.f->evKernel1));
        SPB_CUDA(cudaEventElapsedTime(&totalMs, L.f->evStart, L.f->evEnd));
        int clockKHz = 0;
        cudaDeviceGetAttribute(&clockKHz, cudaDevAttrClockRate, L.device);
        std::vector<unsigned long long> total(CTR_COUNT, 0);
        for (u32 k = 0; k < count; ++k)
        {
            const unsigned long long *ck = c.data() + (size_t)k * CTR_COUNT;
            for (int j = 0; j < CTR_COUNT; ++j) total[j] += ck[j];
            if (metricsBuffer && done + k < maxMetrics)
            {
                // one sp_Metrics per task, in pop order (the reference appends in completion order,
                // main.cpp:746-747); CyclesElapsed = this tile's own device time in nanoseconds
                sp_Metrics m;
                memset(&m, 0, sizeof(m));
                float tileMs = clockKHz > 0 ? (float)((double)ck[CTR_CLOCK_SUM] / (double)clockKHz) : kernelMs;
                add_metrics(&m, ck, tileMs);
                metricsBuffer[done + k] = m;
            }
        }
        record_stats(total.data(), kernelMs, totalMs);
        done += count;
    }
    return done;
}

extern "C" int sp_b200_PrimaryHits(sp_Context *ctx, u32 sample, u32 frame, i32 *triangleIndex,
                                   i32 *objectIndex, f32 *t)
{
    Library &L = lib();
    std::lock_guard<std::recursive_mutex> lock(L.mutex);
    ensure_init();
    SPB_ASSERT(ctx && ctx->camera && ctx->camera->imagePlane);
    DCamera cam;
    convert_camera(ctx->camera, &cam);
    DeviceScene *ds = find_scene(ctx->scene);
    size_t n = (size_t)cam.width * cam.height;
    if (n == 0) return 0;
    L.scratchA.ensure(n * 4);
    L.scratchB.ensure(n * 4);
    L.scratchC.ensure(n * 4);
    unsigned long long *ctr = reset_counters();
    SPB_CUDA(cudaEventRecord(L.f->evStart, L.stream));
    SPB_CUDA(cudaEventRecord(L.f->evKernel0, L.stream));
    launch_primary_hits(kernel_config(), launch_scene(ds), cam, sample, frame, (int32_t *)L.scratchA.ptr,
                        (int32_t *)L.scratchB.ptr, (float *)L.scratchC.ptr, ctr, L.stream);
    SPB_CUDA(cudaGetLastError());
    SPB_CUDA(cudaEventRecord(L.f->evKernel1, L.stream));
    unsigned long long c[CTR_COUNT];
    SPB_CUDA(cudaMemcpyAsync(c, ctr, sizeof(c), cudaMemcpyDeviceToHost, L.stream));
    if (triangleIndex) SPB_CUDA(cudaMemcpyAsync(triangleIndex, L.scratchA.ptr, n * 4, cudaMemcpyDeviceToHost, L.stream));
    if (objectIndex) SPB_CUDA(cudaMemcpyAsync(objectIndex, L.scratchB.ptr, n * 4, cudaMemcpyDeviceToHost, L.stream));
    if (t) SPB_CUDA(cudaMemcpyAsync(t, L.scratchC.ptr, n * 4, cudaMemcpyDeviceToHost, L.stream));
    SPB_CUDA(cudaEventRecord(L.f->evEnd, L.stream));
    SPB_CUDA(cudaStreamSynchronize(L.stream));
    float kernelMs = 0, totalMs = 0;
    SPB_CUDA(cudaEventElapsedTime(&kernelMs, L.f->evKernel0, L.f->evKernel1));
    SPB_CUDA(cudaEventElapsedTime(&totalMs, L.f->evStart, L.f->evEnd));
    record_stats(c, kernelMs, totalMs);
    return 0;
}
