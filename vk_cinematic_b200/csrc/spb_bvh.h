// spb_bvh.h -- host builder of the 4-wide BVH the traversal kernels consume.
//
// Replaces bvh_CreateTree (reference src/bvh.cpp:51-200, an O(n^2 log n) agglomerative build)
// with a binned-SAH top-down build collapsed to 4 children per node.  What is kept from the
// reference is the *predicate structure*, not the topology: every primitive sits alone in a
// child slot whose box is exactly the primitive's own AABB, so the slab test performed at the
// parent is the per-leaf test of bvh_IntersectRay (bvh.cpp:236-255), and a primitive is
// reported iff its own box passes -- independent of how the tree above it is shaped.
#pragma once

#include <stdint.h>
#include <vector>

namespace spb {

struct Node4
{
    float bmin[3][4]; // [axis][lane]
    float bmax[3][4];
    uint32_t ref[4];  // SPB_REF_EMPTY, SPB_REF_LEAF | slot, or node index (local to this tree)
    uint32_t meta[4]; // meta[0] = child count, meta[1] = depth of this node
};
static_assert(sizeof(Node4) == 128, "node must be 128 bytes (8 x 16-byte loads)");

struct Bvh4
{
    std::vector<Node4> nodes;       // breadth-first order: the first nodes are the top levels
    std::vector<uint32_t> slotPrim; // leaf slot -> primitive index (slots in breadth-first order)
    uint32_t maxDepth = 0;
    uint32_t stackNeed = 0; // worst-case traversal stack entries inside this tree
    float rootMin[3] = {0, 0, 0};
    float rootMax[3] = {0, 0, 0};
};

// aabbMin/aabbMax: count x 3 floats.  count == 0 gives an empty tree (no nodes).
// stackLimit: the traversal stack entries this tree may need (0: the share of a mesh tree,
// SPB_MESH_STACK_LIMIT); a SAH tree that would need more is rebuilt with median splits.
Bvh4 build_bvh4(const float *aabbMin, const float *aabbMax, uint32_t count, uint32_t stackLimit = 0);
// Largest |coordinate| of the finite corners of a tree's child boxes (every box of the tree lies
// inside its root's child boxes); 0 for an empty tree.
float bvh4_extent(const Bvh4 &bvh);

// ---- device builder (LBVH, spb_lbvh.cu) --------------------------------------------------------
// The O(n log n) part -- Morton keys, sort, binary radix tree, bottom-up boxes -- runs on the GPU;
// what comes back is the binary tree below.  bvh4_from_binary validates it (every primitive and
// every internal node reached exactly once from node 0), renumbers it parent-before-child and
// runs the same 4-wide collapse as the host builder.  It returns false -- the caller falls back
// to build_bvh4 -- when the tree is malformed or deeper than the traversal stack allows.
struct BinaryTree
{
    std::vector<uint32_t> sortedPrim; // n: primitive index at each sorted position
    std::vector<uint32_t> children;   // 2 x (n - 1): internal node index, or SPB_REF_LEAF | sorted position
    std::vector<float> boxes;         // 6 x (n - 1): min xyz, max xyz of each internal node
};
bool bvh4_from_binary(const float *aabbMin, const float *aabbMax, uint32_t count, const BinaryTree &tree,
                      Bvh4 *out);
// The finished 4-wide tree as the device builder returns it (lbvh_build_bvh4_device): nodes in the layout of
// Node4 (32 words each), numbered level by level -- children after their parents -- but in no particular order
// inside a level (atomic counters); leaf slot -> primitive; what the emission counted.
struct DeviceTree4
{
    std::vector<uint32_t> nodes;    // 32 words per node
    std::vector<uint32_t> slotPrim; // n
    uint32_t maxDepth = 0, stackNeed = 0;
    float rootBox[6] = {0, 0, 0, 0, 0, 0}; // min xyz, max xyz of binary node 0
};
// Checks a DeviceTree4 and turns it into a Bvh4: every primitive in exactly one leaf slot, every node referenced
// exactly once by a node before it, child counts and depths consistent, every node's children inside the box its
// parent holds for it; leaf boxes are OVERWRITTEN with the caller's own AABBs (the parity contract rests on
// those, not on anything the device computed); depth and stack need are recomputed, not trusted.  false --
// the caller falls back to build_bvh4 -- when anything is off or the tree needs more stack than a mesh tree may.
bool bvh4_adopt_device_tree(const float *aabbMin, const float *aabbMax, uint32_t count, const DeviceTree4 &tree,
                            Bvh4 *out);
// Host emulation of the device collapse (the per-element functions of spb_lbvh.cuh run level by level in a loop)
// on the host's binary tree: what tests/hostsim builds with builder 2.
bool lbvh_collapse_host_emulation(const float *aabbMin, const float *aabbMax, uint32_t count, const BinaryTree &tree,
                                  DeviceTree4 *out);
Bvh4 build_bvh4_lbvh_host_device_collapse(const float *aabbMin, const float *aabbMax, uint32_t count);
// The same binary tree computed on the host with the per-element functions the kernels run
// (spb_lbvh.cuh): the reference the device result is compared with, and what tests/hostsim uses.
BinaryTree lbvh_build_binary_host(const float *aabbMin, const float *aabbMax, uint32_t count);
// bounds of all finite box corners (the Morton grid); false if there is no finite extent at all
void lbvh_root_bounds(const float *aabbMin, const float *aabbMax, uint32_t count, float *rootMin, float *rootMax);
// host emulation end to end: LBVH binary tree -> 4-wide, falling back to build_bvh4 like the device path
Bvh4 build_bvh4_lbvh_host(const float *aabbMin, const float *aabbMax, uint32_t count);

} // namespace spb
