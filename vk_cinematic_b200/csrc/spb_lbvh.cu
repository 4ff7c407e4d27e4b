// spb_lbvh.cu -- device side of the LBVH builder (arithmetic in spb_lbvh.cuh,
// host finalisation in spb_bvh.cpp: bvh4_from_binary).
//
// Compile: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -lineinfo.
//
//   k_lbvh_keys   one primitive per thread: 63-bit Morton key of its box centre, value = its index
//   cub::DeviceRadixSort::SortPairs (64-bit keys; the one library call: a radix sort is plumbing)
//   k_lbvh_nodes  one internal node per thread: range and split by binary search over key prefixes
//   k_lbvh_fit    one leaf per thread walks up; the second thread to arrive at a node (one atomic
//                 counter per node) unions the children's boxes and continues -- O(n) work, no
//                 level synchronisation; with `cost` it also runs the node's step of the 4-wide
//                 collapse's dynamic programme (children complete = their costs complete)
//   k_lbvh_emit   (lbvh_build_bvh4_device) one 4-wide node per thread, level by level from the root: gathers
//                 the node's up to four children from the programme's picks, writes the 128-byte node, numbers
//                 children and leaf slots from atomic counters, appends the next level's work items
// All four are HBM-streaming passes over 8-32 bytes per primitive; at the mesh sizes of this
// workload (5 k - 82 k triangles per mesh) they are launch-latency bound, which is the point:
// the reference's builder needs 0.37 s / 4.3 s for bunny / monkey, the host SAH builder 0.4 s for
// the 81 920-triangle sphere.
#include <cub/device/device_radix_sort.cuh>

#include "spb_kernels.cuh"
#include "spb_lbvh.cuh"

namespace spb {

struct LbvhBounds { float mn[3], mx[3]; };

__global__ void __launch_bounds__(256)
k_lbvh_keys(const float *aabbMin, const float *aabbMax, uint32_t count, LbvhBounds root, uint64_t *keys,
            uint32_t *values)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    keys[i] = lbvh_key(aabbMin + (size_t)i * 3, aabbMax + (size_t)i * 3, root.mn, root.mx);
    values[i] = i;
}

__global__ void __launch_bounds__(256)
k_lbvh_nodes(const uint64_t *sortedKeys, uint32_t count, uint32_t *children, uint32_t *internalParent,
             uint32_t *leafParent)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i + 1 >= count) return;
    uint32_t left, right;
    lbvh_node(sortedKeys, count, i, left, right);
    children[(size_t)i * 2] = left;
    children[(size_t)i * 2 + 1] = right;
    if (left & SPB_REF_LEAF) leafParent[left & ~SPB_REF_LEAF] = i; else internalParent[left] = i;
    if (right & SPB_REF_LEAF) leafParent[right & ~SPB_REF_LEAF] = i; else internalParent[right] = i;
    if (i == 0) internalParent[0] = 0xFFFFFFFFu;
}

__device__ __forceinline__ void lbvh_child_box(uint32_t ref, const uint32_t *sortedPrim, const float *aabbMin,
                                               const float *aabbMax, const volatile float *boxes, float *mn,
                                               float *mx)
{
    if (ref & SPB_REF_LEAF)
    {
        uint32_t prim = sortedPrim[ref & ~SPB_REF_LEAF];
        for (int a = 0; a < 3; ++a) { mn[a] = aabbMin[(size_t)prim * 3 + a]; mx[a] = aabbMax[(size_t)prim * 3 + a]; }
    }
    else
    {
        for (int a = 0; a < 3; ++a) { mn[a] = boxes[(size_t)ref * 6 + a]; mx[a] = boxes[(size_t)ref * 6 + 3 + a]; }
    }
}

__global__ void __launch_bounds__(256)
k_lbvh_fit(const uint32_t *sortedPrim, const float *aabbMin, const float *aabbMax, uint32_t count,
           const uint32_t *children, const uint32_t *internalParent, const uint32_t *leafParent,
           uint32_t *arrivals, float *boxes, float *cost, uint32_t *picks)
{
    uint32_t leaf = blockIdx.x * blockDim.x + threadIdx.x;
    if (leaf >= count) return;
    uint32_t node = leafParent[leaf];
    while (node != 0xFFFFFFFFu)
    {
        // the first thread to arrive stops; the second finds both children complete (the first
        // one's box stores are ordered before its atomic by the fence below)
        if (atomicAdd(&arrivals[node], 1u) == 0u) return;
        __threadfence();
        float lmn[3], lmx[3], rmn[3], rmx[3];
        lbvh_child_box(children[(size_t)node * 2], sortedPrim, aabbMin, aabbMax, boxes, lmn, lmx);
        lbvh_child_box(children[(size_t)node * 2 + 1], sortedPrim, aabbMin, aabbMax, boxes, rmn, rmx);
        for (int a = 0; a < 3; ++a)
        {
            boxes[(size_t)node * 6 + a] = fminf(lmn[a], rmn[a]);
            boxes[(size_t)node * 6 + 3 + a] = fmaxf(lmx[a], rmx[a]);
        }
        if (cost)
        {
            // the node's step of the collapse's dynamic programme (spb_lbvh.cuh lbvh_dp)
            const uint32_t l = children[(size_t)node * 2], r = children[(size_t)node * 2 + 1];
            const volatile float *vc = cost;
            float cl[3] = {0.0f, 0.0f, 0.0f}, cr[3] = {0.0f, 0.0f, 0.0f}, box[6], c3[3];
            if (!(l & SPB_REF_LEAF)) for (int i = 0; i < 3; ++i) cl[i] = vc[(size_t)l * 3 + i];
            if (!(r & SPB_REF_LEAF)) for (int i = 0; i < 3; ++i) cr[i] = vc[(size_t)r * 3 + i];
            for (int a = 0; a < 3; ++a) { box[a] = fminf(lmn[a], rmn[a]); box[3 + a] = fmaxf(lmx[a], rmx[a]); }
            uint32_t p;
            lbvh_dp(box, cl, cr, c3, &p);
            for (int i = 0; i < 3; ++i) cost[(size_t)node * 3 + i] = c3[i];
            picks[node] = p;
        }
        __threadfence();
        node = internalParent[node];
    }
}

#define LBVH_CUDA(call)                                                                            \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) { ok = false; goto done; }                                          \
    } while (0)

// Builds the binary radix tree of `count` boxes on the current device.  Returns false on any CUDA
// error (the caller falls back to the host builder); the result is validated by bvh4_from_binary.
bool lbvh_build_binary_device(const float *aabbMin, const float *aabbMax, uint32_t count, BinaryTree *tree,
                              float *kernelMs, cudaStream_t stream)
{
    if (count < 2) return false;
    bool ok = true;
    const uint32_t internal = count - 1;
    LbvhBounds root;
    lbvh_root_bounds(aabbMin, aabbMax, count, root.mn, root.mx);

    // one allocation, carved up (256-byte aligned pieces)
    auto align = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t boxBytes = align((size_t)count * 12);
    size_t sortBytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sortBytes, (const uint64_t *)nullptr, (uint64_t *)nullptr,
                                    (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)count, 0, 63, stream);
    const size_t sizes[] = {boxBytes, boxBytes, align((size_t)count * 8), align((size_t)count * 8),
                            align((size_t)count * 4), align((size_t)count * 4), align((size_t)internal * 8),
                            align((size_t)internal * 4), align((size_t)count * 4), align((size_t)internal * 4),
                            align((size_t)internal * 24), align(sortBytes)};
    size_t total = 0;
    for (size_t s : sizes) total += s;
    char *base = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    g_kernelLaunches += 4;
    {
        LBVH_CUDA(cudaMalloc(&base, total));
        char *at = base;
        auto take = [&](size_t bytes) { char *p = at; at += bytes; return p; };
        float *dMin = (float *)take(sizes[0]);
        float *dMax = (float *)take(sizes[1]);
        uint64_t *keysIn = (uint64_t *)take(sizes[2]);
        uint64_t *keysOut = (uint64_t *)take(sizes[3]);
        uint32_t *valuesIn = (uint32_t *)take(sizes[4]);
        uint32_t *valuesOut = (uint32_t *)take(sizes[5]);
        uint32_t *children = (uint32_t *)take(sizes[6]);
        uint32_t *internalParent = (uint32_t *)take(sizes[7]);
        uint32_t *leafParent = (uint32_t *)take(sizes[8]);
        uint32_t *arrivals = (uint32_t *)take(sizes[9]);
        float *boxes = (float *)take(sizes[10]);
        void *sortTemp = take(sizes[11]);

        LBVH_CUDA(cudaEventCreate(&e0));
        LBVH_CUDA(cudaEventCreate(&e1));
        LBVH_CUDA(cudaMemcpyAsync(dMin, aabbMin, (size_t)count * 12, cudaMemcpyHostToDevice, stream));
        LBVH_CUDA(cudaMemcpyAsync(dMax, aabbMax, (size_t)count * 12, cudaMemcpyHostToDevice, stream));
        LBVH_CUDA(cudaMemsetAsync(arrivals, 0, (size_t)internal * 4, stream));
        LBVH_CUDA(cudaEventRecord(e0, stream));
        const unsigned blocks = (count + 255u) / 256u;
        k_lbvh_keys<<<blocks, 256, 0, stream>>>(dMin, dMax, count, root, keysIn, valuesIn);
        LBVH_CUDA(cudaGetLastError());
        LBVH_CUDA(cub::DeviceRadixSort::SortPairs(sortTemp, sortBytes, keysIn, keysOut, valuesIn, valuesOut, (int)count,
                                                  0, 63, stream));
        k_lbvh_nodes<<<blocks, 256, 0, stream>>>(keysOut, count, children, internalParent, leafParent);
        LBVH_CUDA(cudaGetLastError());
        k_lbvh_fit<<<blocks, 256, 0, stream>>>(valuesOut, dMin, dMax, count, children, internalParent, leafParent,
                                               arrivals, boxes, nullptr, nullptr);
        LBVH_CUDA(cudaGetLastError());
        LBVH_CUDA(cudaEventRecord(e1, stream));

        tree->sortedPrim.resize(count);
        tree->children.resize((size_t)internal * 2);
        tree->boxes.resize((size_t)internal * 6);
        LBVH_CUDA(cudaMemcpyAsync(tree->sortedPrim.data(), valuesOut, (size_t)count * 4, cudaMemcpyDeviceToHost, stream));
        LBVH_CUDA(cudaMemcpyAsync(tree->children.data(), children, (size_t)internal * 8, cudaMemcpyDeviceToHost, stream));
        LBVH_CUDA(cudaMemcpyAsync(tree->boxes.data(), boxes, (size_t)internal * 24, cudaMemcpyDeviceToHost, stream));
        LBVH_CUDA(cudaStreamSynchronize(stream));
        if (kernelMs) LBVH_CUDA(cudaEventElapsedTime(kernelMs, e0, e1));
    }
done:
    if (!ok) cudaGetLastError(); // clear: the caller falls back to the host builder
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (base) cudaFree(base);
    return ok;
}

__global__ void k_lbvh_seed(uint32_t *counters, uint32_t *levelCount)
{
    if (threadIdx.x == 0) { counters[0] = 1u; levelCount[0] = 1u; }
}

// One level of the emission.  The number of items of a level is known on the device only (levelCount[level],
// counted by the level before): the launch covers the most a level can hold and the surplus threads leave, so the
// host never waits between levels.
__global__ void __launch_bounds__(128)
k_lbvh_emit(LbvhDpView view, const uint32_t *sortedPrim, const float *primMin, const float *primMax, const float *boxes,
            const LbvhEmitItem *items, const uint32_t *levelCount, uint32_t *nodes4, uint32_t *slotPrim, uint32_t *counters,
            LbvhEmitItem *next, uint32_t *nextCount)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *levelCount) return;
    lbvh_emit(view, sortedPrim, primMin, primMax, boxes, items[i], nodes4, slotPrim, counters, next, nextCount);
}

// The whole build on the device: keys, sort, radix tree, boxes + dynamic programme, level-by-level emission of
// the 4-wide nodes.  What comes back is the finished tree (128 bytes per node, the leaf-slot permutation, five
// counters); the host checks it (bvh4_adopt_device_tree, spb_bvh.cpp) instead of building it.  One wait for the
// device, at the end: the levels are launched back to back for the deepest tree a mesh may have
// (SPB_MESH_STACK_LIMIT entries = at least one per level), each over the most items a level can hold, and a level
// past the last finds a count of 0.  Device memory comes from the stream-ordered pool (no cudaMalloc / cudaFree,
// which wait for the whole device).  Returns false on any CUDA error or a tree with levels left over; the caller
// falls back to the host builder.
bool lbvh_build_bvh4_device(const float *aabbMin, const float *aabbMax, uint32_t count, DeviceTree4 *tree,
                            float *kernelMs, cudaStream_t stream)
{
    if (count < 2) return false;
    bool ok = true;
    const uint32_t internal = count - 1, levels = SPB_MESH_STACK_LIMIT + 2u;
    LbvhBounds root;
    lbvh_root_bounds(aabbMin, aabbMax, count, root.mn, root.mx);
    auto align = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t boxBytes = align((size_t)count * 12);
    size_t sortBytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sortBytes, (const uint64_t *)nullptr, (uint64_t *)nullptr,
                                    (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)count, 0, 63, stream);
    const size_t sizes[] = {boxBytes, boxBytes, align((size_t)count * 8), align((size_t)count * 8),
                            align((size_t)count * 4), align((size_t)count * 4), align((size_t)internal * 8),
                            align((size_t)internal * 4), align((size_t)count * 4), align((size_t)internal * 4),
                            align((size_t)internal * 24), align(sortBytes),
                            align((size_t)internal * 12), align((size_t)internal * 4),              // cost, picks
                            align((size_t)internal * 128), align((size_t)count * 4),               // nodes4, slotPrim
                            align((size_t)internal * 16), align((size_t)internal * 16),            // items x 2
                            align(((size_t)levels + 1 + 8) * 4)};                                  // counters[8], levelCount[levels + 1]
    size_t total = 0;
    for (size_t s : sizes) total += s;
    char *base = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    std::vector<uint32_t> tail((size_t)levels + 1 + 8, 0u);
    {
        LBVH_CUDA(cudaMallocAsync((void **)&base, total, stream));
        char *at = base;
        auto take = [&](size_t bytes) { char *p = at; at += bytes; return p; };
        float *dMin = (float *)take(sizes[0]);
        float *dMax = (float *)take(sizes[1]);
        uint64_t *keysIn = (uint64_t *)take(sizes[2]);
        uint64_t *keysOut = (uint64_t *)take(sizes[3]);
        uint32_t *valuesIn = (uint32_t *)take(sizes[4]);
        uint32_t *valuesOut = (uint32_t *)take(sizes[5]);
        uint32_t *children = (uint32_t *)take(sizes[6]);
        uint32_t *internalParent = (uint32_t *)take(sizes[7]);
        uint32_t *leafParent = (uint32_t *)take(sizes[8]);
        uint32_t *arrivals = (uint32_t *)take(sizes[9]);
        float *boxes = (float *)take(sizes[10]);
        void *sortTemp = take(sizes[11]);
        float *cost = (float *)take(sizes[12]);
        uint32_t *picks = (uint32_t *)take(sizes[13]);
        uint32_t *nodes4 = (uint32_t *)take(sizes[14]);
        uint32_t *slotPrim = (uint32_t *)take(sizes[15]);
        LbvhEmitItem *items[2] = {(LbvhEmitItem *)take(sizes[16]), (LbvhEmitItem *)take(sizes[17])};
        uint32_t *counters = (uint32_t *)take(sizes[18]);
        uint32_t *levelCount = counters + 8;

        LBVH_CUDA(cudaEventCreate(&e0));
        LBVH_CUDA(cudaEventCreate(&e1));
        LBVH_CUDA(cudaMemcpyAsync(dMin, aabbMin, (size_t)count * 12, cudaMemcpyHostToDevice, stream));
        LBVH_CUDA(cudaMemcpyAsync(dMax, aabbMax, (size_t)count * 12, cudaMemcpyHostToDevice, stream));
        LBVH_CUDA(cudaMemsetAsync(arrivals, 0, (size_t)internal * 4, stream));
        // counters = {nodes: the root, leaf slots, -, deepest node, worst stack}; levelCount[0] = 1 item: the root
        LBVH_CUDA(cudaMemsetAsync(counters, 0, ((size_t)levels + 1 + 8) * 4, stream));
        LBVH_CUDA(cudaMemsetAsync(items[0], 0, sizeof(LbvhEmitItem), stream)); // (binary node 0 -> 4-wide node 0, depth 0)
        LBVH_CUDA(cudaEventRecord(e0, stream));
        const unsigned blocks = (count + 255u) / 256u;
        g_kernelLaunches += 5;
        k_lbvh_keys<<<blocks, 256, 0, stream>>>(dMin, dMax, count, root, keysIn, valuesIn);
        LBVH_CUDA(cudaGetLastError());
        LBVH_CUDA(cub::DeviceRadixSort::SortPairs(sortTemp, sortBytes, keysIn, keysOut, valuesIn, valuesOut, (int)count,
                                                  0, 63, stream));
        k_lbvh_nodes<<<blocks, 256, 0, stream>>>(keysOut, count, children, internalParent, leafParent);
        LBVH_CUDA(cudaGetLastError());
        k_lbvh_fit<<<blocks, 256, 0, stream>>>(valuesOut, dMin, dMax, count, children, internalParent, leafParent,
                                               arrivals, boxes, cost, picks);
        LBVH_CUDA(cudaGetLastError());
        k_lbvh_seed<<<1, 32, 0, stream>>>(counters, levelCount);
        LBVH_CUDA(cudaGetLastError());
        LbvhDpView view = {children, cost, picks};
        uint64_t most = 1; // items a level can hold: 4 per item of the level before, never more than there are internal nodes
        for (uint32_t level = 0; level < levels; ++level)
        {
            const uint32_t threads = (uint32_t)(most < internal ? most : internal);
            k_lbvh_emit<<<(threads + 127u) / 128u, 128, 0, stream>>>(view, valuesOut, dMin, dMax, boxes, items[level & 1u], levelCount + level,
                                                                    nodes4, slotPrim, counters, items[(level + 1u) & 1u], levelCount + level + 1);
            g_kernelLaunches++;
            if (most < internal) most *= 4;
        }
        LBVH_CUDA(cudaGetLastError());
        LBVH_CUDA(cudaEventRecord(e1, stream));
        LBVH_CUDA(cudaMemcpyAsync(tail.data(), counters, tail.size() * 4, cudaMemcpyDeviceToHost, stream));
        LBVH_CUDA(cudaMemcpyAsync(tree->rootBox, boxes, 24, cudaMemcpyDeviceToHost, stream));
        LBVH_CUDA(cudaStreamSynchronize(stream));
        const uint32_t nodeCount = tail[0], slots = tail[1];
        // (a tree with levels left over, or counts that cannot be: refused here, before anything is sized by them)
        if (tail[8 + levels] != 0 || nodeCount < 1 || nodeCount > internal || slots != count) { ok = false; goto done; }
        tree->nodes.resize((size_t)nodeCount * 32);
        tree->slotPrim.resize(count);
        tree->maxDepth = tail[3];
        tree->stackNeed = tail[4];
        LBVH_CUDA(cudaMemcpyAsync(tree->nodes.data(), nodes4, (size_t)nodeCount * 128, cudaMemcpyDeviceToHost, stream));
        LBVH_CUDA(cudaMemcpyAsync(tree->slotPrim.data(), slotPrim, (size_t)count * 4, cudaMemcpyDeviceToHost, stream));
        LBVH_CUDA(cudaStreamSynchronize(stream));
        if (kernelMs) LBVH_CUDA(cudaEventElapsedTime(kernelMs, e0, e1));
    }
done:
    if (!ok) cudaGetLastError();
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (base) cudaFreeAsync(base, stream);
    return ok;
}

} // namespace spb
