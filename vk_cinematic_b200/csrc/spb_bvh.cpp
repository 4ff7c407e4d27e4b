// spb_bvh.cpp -- binned-SAH binary build, then a greedy collapse to 4 children per node.
// See spb_bvh.h for what is and is not kept from the reference's bvh_CreateTree.
#include "spb_bvh.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <queue>

#include "spb_core.cuh"
#include "spb_lbvh.cuh"

#ifndef SPB_SWEEP_LIMIT
#define SPB_SWEEP_LIMIT 8192u
#endif

namespace spb {

namespace {

struct BNode
{
    float mn[3], mx[3];
    uint32_t left, right;  // children (internal) ...
    uint32_t first, count; // ... or the range of `order` covered (count == 1 -> leaf)
};

inline float half_area(const float *mn, const float *mx)
{
    float dx = mx[0] - mn[0], dy = mx[1] - mn[1], dz = mx[2] - mn[2];
    return dx * dy + dy * dz + dz * dx;
}

struct Builder
{
    const float *aabbMin;
    const float *aabbMax;
    std::vector<uint32_t> order;
    std::vector<float> centroid; // 3 per primitive
    std::vector<BNode> nodes;
    bool balancedOnly;

    void bounds_of(uint32_t first, uint32_t count, float *mn, float *mx) const
    {
        for (int a = 0; a < 3; ++a) { mn[a] = INFINITY; mx[a] = -INFINITY; }
        for (uint32_t i = first; i < first + count; ++i)
        {
            const float *pmn = aabbMin + (size_t)order[i] * 3;
            const float *pmx = aabbMax + (size_t)order[i] * 3;
            for (int a = 0; a < 3; ++a)
            {
                // NaN-free exact min/max; a NaN box keeps its NaN out of the ancestors
                if (pmn[a] < mn[a]) mn[a] = pmn[a];
                if (pmx[a] > mx[a]) mx[a] = pmx[a];
            }
        }
    }

    // Exact SAH for small ranges: sort by centroid on each axis and sweep every split position.
    // Returns the split position or 0 when no axis separates the centroids.
    uint32_t sweep_split(uint32_t first, uint32_t count)
    {
        std::vector<uint32_t> best, cur(order.begin() + first, order.begin() + first + count);
        std::vector<float> rightArea(count);
        float bestCost = INFINITY;
        uint32_t bestPos = 0;
        for (int axis = 0; axis < 3; ++axis)
        {
            std::sort(cur.begin(), cur.end(), [&](uint32_t a, uint32_t b) {
                float ca = centroid[(size_t)a * 3 + axis], cb = centroid[(size_t)b * 3 + axis];
                return ca < cb || (ca == cb && a < b);
            });
            if (!(centroid[(size_t)cur.front() * 3 + axis] < centroid[(size_t)cur.back() * 3 + axis])) continue;
            float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
            for (uint32_t i = count - 1; i > 0; --i)
            {
                for (int a = 0; a < 3; ++a)
                {
                    mn[a] = std::min(mn[a], aabbMin[(size_t)cur[i] * 3 + a]);
                    mx[a] = std::max(mx[a], aabbMax[(size_t)cur[i] * 3 + a]);
                }
                rightArea[i] = half_area(mn, mx);
            }
            for (int a = 0; a < 3; ++a) { mn[a] = INFINITY; mx[a] = -INFINITY; }
            bool improved = false;
            for (uint32_t i = 1; i < count; ++i)
            {
                for (int a = 0; a < 3; ++a)
                {
                    mn[a] = std::min(mn[a], aabbMin[(size_t)cur[i - 1] * 3 + a]);
                    mx[a] = std::max(mx[a], aabbMax[(size_t)cur[i - 1] * 3 + a]);
                }
                float cost = half_area(mn, mx) * (float)i + rightArea[i] * (float)(count - i);
                if (cost < bestCost)
                {
                    bestCost = cost;
                    bestPos = i;
                    improved = true;
                }
            }
            if (improved) best = cur;
        }
        if (bestPos == 0) return 0;
        std::copy(best.begin(), best.end(), order.begin() + first);
        return first + bestPos;
    }

    // Returns the split position (first index of the right half) after partitioning `order`.
    uint32_t split(uint32_t first, uint32_t count, const float *nodeMin, const float *nodeMax)
    {
        const int BINS = 16;
        if (!balancedOnly && count > 2 && count <= SPB_SWEEP_LIMIT)
        {
            uint32_t pos = sweep_split(first, count);
            if (pos) return pos;
        }
        uint32_t mid = first + count / 2;
        float cmn[3] = {INFINITY, INFINITY, INFINITY}, cmx[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (uint32_t i = first; i < first + count; ++i)
        {
            const float *c = &centroid[(size_t)order[i] * 3];
            for (int a = 0; a < 3; ++a)
            {
                if (c[a] < cmn[a]) cmn[a] = c[a];
                if (c[a] > cmx[a]) cmx[a] = c[a];
            }
        }
        int bestAxis = -1, bestBin = -1;
        float bestCost = INFINITY;
        if (!balancedOnly && count > 2)
        {
            for (int axis = 0; axis < 3; ++axis)
            {
                float extent = cmx[axis] - cmn[axis];
                if (!(extent > 0.0f) || !std::isfinite(extent)) continue;
                float scale = (float)BINS / extent;
                uint32_t binCount[BINS] = {};
                float binMin[BINS][3], binMax[BINS][3];
                for (int b = 0; b < BINS; ++b)
                    for (int a = 0; a < 3; ++a) { binMin[b][a] = INFINITY; binMax[b][a] = -INFINITY; }
                for (uint32_t i = first; i < first + count; ++i)
                {
                    uint32_t p = order[i];
                    int b = (int)((centroid[(size_t)p * 3 + axis] - cmn[axis]) * scale);
                    b = b < 0 ? 0 : (b >= BINS ? BINS - 1 : b);
                    binCount[b]++;
                    for (int a = 0; a < 3; ++a)
                    {
                        binMin[b][a] = std::min(binMin[b][a], aabbMin[(size_t)p * 3 + a]);
                        binMax[b][a] = std::max(binMax[b][a], aabbMax[(size_t)p * 3 + a]);
                    }
                }
                float rightArea[BINS];
                uint32_t rightCount[BINS];
                float rmn[3] = {INFINITY, INFINITY, INFINITY}, rmx[3] = {-INFINITY, -INFINITY, -INFINITY};
                uint32_t rc = 0;
                for (int b = BINS - 1; b > 0; --b)
                {
                    for (int a = 0; a < 3; ++a)
                    {
                        rmn[a] = std::min(rmn[a], binMin[b][a]);
                        rmx[a] = std::max(rmx[a], binMax[b][a]);
                    }
                    rc += binCount[b];
                    rightCount[b] = rc;
                    rightArea[b] = rc ? half_area(rmn, rmx) : 0.0f;
                }
                float lmn[3] = {INFINITY, INFINITY, INFINITY}, lmx[3] = {-INFINITY, -INFINITY, -INFINITY};
                uint32_t lc = 0;
                for (int b = 0; b < BINS - 1; ++b)
                {
                    for (int a = 0; a < 3; ++a)
                    {
                        lmn[a] = std::min(lmn[a], binMin[b][a]);
                        lmx[a] = std::max(lmx[a], binMax[b][a]);
                    }
                    lc += binCount[b];
                    if (lc == 0 || rightCount[b + 1] == 0) continue;
                    float cost = half_area(lmn, lmx) * (float)lc + rightArea[b + 1] * (float)rightCount[b + 1];
                    if (cost < bestCost)
                    {
                        bestCost = cost;
                        bestAxis = axis;
                        bestBin = b;
                    }
                }
            }
        }
        if (bestAxis >= 0)
        {
            float extent = cmx[bestAxis] - cmn[bestAxis];
            float scale = (float)BINS / extent;
            float lo = cmn[bestAxis];
            int axis = bestAxis, bin = bestBin;
            auto it = std::partition(order.begin() + first, order.begin() + first + count,
                [&](uint32_t p) {
                    int b = (int)((centroid[(size_t)p * 3 + axis] - lo) * scale);
                    b = b < 0 ? 0 : (b >= BINS ? BINS - 1 : b);
                    return b <= bin;
                });
            uint32_t pos = (uint32_t)(it - order.begin());
            if (pos > first && pos < first + count) return pos;
        }
        // median split along the widest centroid axis (also the balanced-only mode)
        int axis = 0;
        float ext[3] = {cmx[0] - cmn[0], cmx[1] - cmn[1], cmx[2] - cmn[2]};
        if (ext[1] > ext[axis]) axis = 1;
        if (ext[2] > ext[axis]) axis = 2;
        std::nth_element(order.begin() + first, order.begin() + mid, order.begin() + first + count,
            [&](uint32_t a, uint32_t b) {
                float ca = centroid[(size_t)a * 3 + axis], cb = centroid[(size_t)b * 3 + axis];
                return ca < cb || (ca == cb && a < b);
            });
        (void)nodeMin;
        (void)nodeMax;
        return mid;
    }

    void build(uint32_t count)
    {
        order.resize(count);
        for (uint32_t i = 0; i < count; ++i) order[i] = i;
        centroid.resize((size_t)count * 3);
        for (size_t i = 0; i < (size_t)count * 3; ++i) centroid[i] = (aabbMin[i] + aabbMax[i]) * 0.5f;
        nodes.clear();
        nodes.reserve((size_t)count * 2);
        BNode root = {};
        root.first = 0;
        root.count = count;
        bounds_of(0, count, root.mn, root.mx);
        nodes.push_back(root);
        std::vector<uint32_t> work;
        work.push_back(0);
        while (!work.empty())
        {
            uint32_t ni = work.back();
            work.pop_back();
            uint32_t first = nodes[ni].first, cnt = nodes[ni].count;
            if (cnt <= 1) continue;
            uint32_t pos = split(first, cnt, nodes[ni].mn, nodes[ni].mx);
            BNode l = {}, r = {};
            l.first = first;
            l.count = pos - first;
            r.first = pos;
            r.count = first + cnt - pos;
            bounds_of(l.first, l.count, l.mn, l.mx);
            bounds_of(r.first, r.count, r.mn, r.mx);
            uint32_t li = (uint32_t)nodes.size();
            nodes.push_back(l);
            nodes.push_back(r);
            nodes[ni].left = li;
            nodes[ni].right = li + 1;
            work.push_back(li);
            work.push_back(li + 1);
        }
    }
};

// Which binary subtrees become the (up to four) children of each 4-wide node, chosen to minimise
// the summed surface area of the 4-wide nodes (the expected number of node visits of a random
// ray) by dynamic programming over the binary tree -- the construction of Ylitie, Karras and
// Laine, "Efficient incoherent ray traversal on GPUs through compressed wide BVHs" (HPG 2017),
// section 4.1, for width 4 and single-primitive leaves:
//   C(n, 1) = area(n) + min over k of C(left, k) + C(right, 4 - k)      n roots a 4-wide node
//   C(n, i) = min(C(n, i - 1), min over k of C(left, k) + C(right, i - k))   n is dissolved into i roots
//   C(leaf, i) = 0
struct WideDp
{
    const Builder &b;
    std::vector<float> cost;    // [node][i - 1], i = 1..3
    std::vector<uint8_t> pick;  // [node][j - 2], j = 2..4: k of the best distribution for j roots
    explicit WideDp(const Builder &builder) : b(builder)
    {
        size_t n = b.nodes.size();
        cost.assign(n * 3, 0.0f);
        pick.assign(n * 3, 1);
        // children are created after their parent (nodes.push_back in build()): reverse index
        // order is a post-order
        for (size_t idx = n; idx-- > 0;)
        {
            const BNode &nd = b.nodes[idx];
            if (nd.count <= 1) continue;
            const float *cl = &cost[(size_t)nd.left * 3], *cr = &cost[(size_t)nd.right * 3];
            float dist[5];
            for (int j = 2; j <= 4; ++j)
            {
                float best = INFINITY;
                int bestK = 1;
                for (int k = 1; k < j; ++k)
                {
                    if (k > 3 || j - k > 3) continue;
                    float c = cl[k - 1] + cr[j - k - 1];
                    if (c < best) { best = c; bestK = k; }
                }
                dist[j] = best;
                pick[idx * 3 + (j - 2)] = (uint8_t)bestK;
            }
            float a = half_area(nd.mn, nd.mx);
            if (!(a >= 0.0f)) a = 0.0f;
            cost[idx * 3 + 0] = a + dist[4];
            cost[idx * 3 + 1] = std::min(dist[2], cost[idx * 3 + 0]);
            cost[idx * 3 + 2] = std::min(dist[3], cost[idx * 3 + 1]);
        }
    }
    // the roots that represent subtree `n` when it may use up to `j` child slots
    void gather(uint32_t n, int j, uint32_t *kids, uint32_t &count) const
    {
        const BNode &nd = b.nodes[n];
        if (nd.count <= 1 || j <= 1) { kids[count++] = n; return; }
        if (j <= 3)
        {
            // C(n, j) = min(dist[j], C(n, j - 1)): prefer fewer roots on ties
            float distJ = cost[(size_t)nd.left * 3 + pick[(size_t)n * 3 + (j - 2)] - 1] +
                          cost[(size_t)nd.right * 3 + (j - pick[(size_t)n * 3 + (j - 2)]) - 1];
            if (!(distJ < cost[(size_t)n * 3 + (j - 2)])) { gather(n, j - 1, kids, count); return; }
        }
        int k = pick[(size_t)n * 3 + (j - 2)];
        gather(nd.left, k, kids, count);
        gather(nd.right, j - k, kids, count);
    }
};

// Collapse in breadth-first order.  Returns the worst-case traversal stack depth.
uint32_t collapse(const Builder &b, Bvh4 &out)
{
    WideDp dp(b);
    out.nodes.clear();
    out.slotPrim.clear();
    out.maxDepth = 0;
    struct Item { uint32_t bnode; uint32_t node4; uint32_t depth; uint32_t stackBefore; };
    std::queue<Item> q;
    out.nodes.emplace_back();
    q.push({0, 0, 0, 0});
    uint32_t worstStack = 0;
    while (!q.empty())
    {
        Item it = q.front();
        q.pop();
        const BNode &bn = b.nodes[it.bnode];
        uint32_t kids[4];
        uint32_t n = 0;
        if (bn.count <= 1)
        {
            kids[n++] = it.bnode; // single-primitive tree: the root node holds one leaf child
        }
        else
        {
#if defined(SPB_GREEDY_COLLAPSE)
            kids[n++] = bn.left;
            kids[n++] = bn.right;
            while (n < 4)
            {
                int pick = -1;
                float area = -1.0f;
                for (uint32_t k = 0; k < n; ++k)
                {
                    const BNode &c = b.nodes[kids[k]];
                    if (c.count <= 1) continue;
                    float a = half_area(c.mn, c.mx);
                    if (!(a >= 0.0f)) a = 0.0f;
                    if (a > area) { area = a; pick = (int)k; }
                }
                if (pick < 0) break;
                uint32_t expand = kids[pick];
                kids[pick] = b.nodes[expand].left;
                kids[n++] = b.nodes[expand].right;
            }
#else
            int k = dp.pick[(size_t)it.bnode * 3 + 2];
            dp.gather(bn.left, k, kids, n);
            dp.gather(bn.right, 4 - k, kids, n);
#endif
        }
        Node4 node;
        memset(&node, 0, sizeof(node));
        uint32_t internalKids = 0;
        for (uint32_t k = 0; k < 4; ++k)
        {
            if (k >= n)
            {
                // empty lane: NaN box never passes either form of the slab test
                for (int a = 0; a < 3; ++a) { node.bmin[a][k] = NAN; node.bmax[a][k] = NAN; }
                node.ref[k] = SPB_REF_EMPTY;
                continue;
            }
            const BNode &c = b.nodes[kids[k]];
            for (int a = 0; a < 3; ++a) { node.bmin[a][k] = c.mn[a]; node.bmax[a][k] = c.mx[a]; }
            if (c.count <= 1)
            {
                uint32_t slot = (uint32_t)out.slotPrim.size();
                out.slotPrim.push_back(b.order[c.first]);
                node.ref[k] = SPB_REF_LEAF | slot;
            }
            else
            {
                internalKids++;
            }
        }
        node.meta[0] = n;
        node.meta[1] = it.depth;
        // the resumable traversal (spb_core.cuh trav_node) pushes every child it does not continue
        // with, primitives included: up to n - 1 entries per level
        (void)internalKids;
        uint32_t stackHere = it.stackBefore + (n > 0 ? n - 1 : 0);
        if (stackHere > worstStack) worstStack = stackHere;
        for (uint32_t k = 0; k < n; ++k)
        {
            const BNode &c = b.nodes[kids[k]];
            if (c.count <= 1) continue;
            uint32_t child4 = (uint32_t)out.nodes.size();
            out.nodes.emplace_back();
            node.ref[k] = child4;
            q.push({kids[k], child4, it.depth + 1, stackHere});
        }
        out.nodes[it.node4] = node;
        if (it.depth > out.maxDepth) out.maxDepth = it.depth;
    }
    return worstStack;
}

} // namespace

void lbvh_root_bounds(const float *aabbMin, const float *aabbMax, uint32_t count, float *rootMin, float *rootMax)
{
    for (int a = 0; a < 3; ++a) { rootMin[a] = INFINITY; rootMax[a] = -INFINITY; }
    for (uint32_t i = 0; i < count; ++i)
        for (int a = 0; a < 3; ++a)
        {
            float lo = aabbMin[(size_t)i * 3 + a], hi = aabbMax[(size_t)i * 3 + a];
            if (std::isfinite(lo) && lo < rootMin[a]) rootMin[a] = lo;
            if (std::isfinite(hi) && hi > rootMax[a]) rootMax[a] = hi;
        }
}

BinaryTree lbvh_build_binary_host(const float *aabbMin, const float *aabbMax, uint32_t count)
{
    BinaryTree t;
    if (count < 2) return t;
    float rootMin[3], rootMax[3];
    lbvh_root_bounds(aabbMin, aabbMax, count, rootMin, rootMax);
    std::vector<uint64_t> keys(count);
    t.sortedPrim.resize(count);
    for (uint32_t i = 0; i < count; ++i)
    {
        keys[i] = lbvh_key(aabbMin + (size_t)i * 3, aabbMax + (size_t)i * 3, rootMin, rootMax);
        t.sortedPrim[i] = i;
    }
    // a radix sort is stable: equal keys keep their input (= primitive index) order
    std::stable_sort(t.sortedPrim.begin(), t.sortedPrim.end(),
                     [&](uint32_t a, uint32_t b) { return keys[a] < keys[b]; });
    std::vector<uint64_t> sortedKeys(count);
    for (uint32_t i = 0; i < count; ++i) sortedKeys[i] = keys[t.sortedPrim[i]];
    const uint32_t internal = count - 1;
    t.children.resize((size_t)internal * 2);
    for (uint32_t i = 0; i < internal; ++i)
        lbvh_node(sortedKeys.data(), count, i, t.children[(size_t)i * 2], t.children[(size_t)i * 2 + 1]);
    // boxes bottom-up: the kernels do this with one atomic counter per node, here a post-order walk
    t.boxes.assign((size_t)internal * 6, 0.0f);
    std::vector<uint8_t> state(internal, 0);
    std::vector<uint32_t> stack;
    stack.push_back(0);
    auto child_box = [&](uint32_t ref, float *mn, float *mx) {
        if (ref & SPB_REF_LEAF)
        {
            uint32_t prim = t.sortedPrim[ref & ~SPB_REF_LEAF];
            for (int a = 0; a < 3; ++a) { mn[a] = aabbMin[(size_t)prim * 3 + a]; mx[a] = aabbMax[(size_t)prim * 3 + a]; }
        }
        else
            for (int a = 0; a < 3; ++a) { mn[a] = t.boxes[(size_t)ref * 6 + a]; mx[a] = t.boxes[(size_t)ref * 6 + 3 + a]; }
    };
    while (!stack.empty())
    {
        uint32_t n = stack.back();
        if (n >= internal) { t.children.clear(); return t; } // malformed: caught by the caller
        if (state[n] == 0)
        {
            state[n] = 1;
            for (int k = 0; k < 2; ++k)
            {
                uint32_t ref = t.children[(size_t)n * 2 + k];
                if (!(ref & SPB_REF_LEAF) && ref < internal && state[ref] == 0) stack.push_back(ref);
            }
            continue;
        }
        stack.pop_back();
        if (state[n] == 2) continue;
        state[n] = 2;
        float lmn[3], lmx[3], rmn[3], rmx[3];
        child_box(t.children[(size_t)n * 2], lmn, lmx);
        child_box(t.children[(size_t)n * 2 + 1], rmn, rmx);
        for (int a = 0; a < 3; ++a)
        {
            // fminf / fmaxf: a NaN operand is ignored, like the `<` / `>` tests of bounds_of
            t.boxes[(size_t)n * 6 + a] = fminf(lmn[a], rmn[a]);
            t.boxes[(size_t)n * 6 + 3 + a] = fmaxf(lmx[a], rmx[a]);
        }
    }
    return t;
}

bool bvh4_from_binary(const float *aabbMin, const float *aabbMax, uint32_t count, const BinaryTree &tree,
                      Bvh4 *out)
{
    if (count < 2) return false;
    const uint32_t internal = count - 1;
    if (tree.sortedPrim.size() != count || tree.children.size() != (size_t)internal * 2 ||
        tree.boxes.size() != (size_t)internal * 6)
        return false;
    // sortedPrim must be a permutation
    {
        std::vector<uint8_t> seen(count, 0);
        for (uint32_t p : tree.sortedPrim)
        {
            if (p >= count || seen[p]) return false;
            seen[p] = 1;
        }
    }
    Builder b;
    b.aabbMin = aabbMin;
    b.aabbMax = aabbMax;
    b.balancedOnly = false;
    b.order = tree.sortedPrim;
    b.nodes.reserve((size_t)count * 2);
    // breadth-first renumbering from internal node 0: parents before children, which is what the
    // collapse's dynamic programme (reverse index order = post-order) relies on
    std::vector<uint8_t> visitedInternal(internal, 0), visitedLeaf(count, 0);
    std::vector<uint32_t> source; // b.nodes index -> tree reference
    auto make = [&](uint32_t ref) -> bool {
        BNode n = {};
        if (ref & SPB_REF_LEAF)
        {
            uint32_t pos = ref & ~SPB_REF_LEAF;
            if (pos >= count || visitedLeaf[pos]) return false;
            visitedLeaf[pos] = 1;
            uint32_t prim = tree.sortedPrim[pos];
            n.first = pos;
            n.count = 1;
            // a leaf's box is the primitive's own AABB, taken from the caller's arrays: the parity
            // contract rests on this, not on anything the device computed
            for (int a = 0; a < 3; ++a) { n.mn[a] = aabbMin[(size_t)prim * 3 + a]; n.mx[a] = aabbMax[(size_t)prim * 3 + a]; }
        }
        else
        {
            if (ref >= internal || visitedInternal[ref]) return false;
            visitedInternal[ref] = 1;
            n.count = 2; // "more than one": the collapse only tests count <= 1
            for (int a = 0; a < 3; ++a) { n.mn[a] = tree.boxes[(size_t)ref * 6 + a]; n.mx[a] = tree.boxes[(size_t)ref * 6 + 3 + a]; }
        }
        b.nodes.push_back(n);
        source.push_back(ref);
        return true;
    };
    if (!make(0)) return false;
    for (size_t at = 0; at < b.nodes.size(); ++at)
    {
        if (b.nodes[at].count <= 1) continue;
        uint32_t ref = source[at];
        uint32_t li = (uint32_t)b.nodes.size();
        if (!make(tree.children[(size_t)ref * 2]) || !make(tree.children[(size_t)ref * 2 + 1])) return false;
        b.nodes[at].left = li;
        b.nodes[at].right = li + 1;
    }
    if (b.nodes.size() != (size_t)count * 2 - 1) return false; // something was never reached
    // every internal box must contain its children's (NaN boxes aside): a cheap check of the fit pass
    for (const BNode &n : b.nodes)
    {
        if (n.count <= 1) continue;
        for (uint32_t child : {n.left, n.right})
            for (int a = 0; a < 3; ++a)
            {
                if (b.nodes[child].mn[a] < n.mn[a] || b.nodes[child].mx[a] > n.mx[a]) return false;
            }
    }
    Bvh4 result;
    uint32_t need = collapse(b, result);
    if (need > SPB_MESH_STACK_LIMIT) return false; // deeper than the traversal stack allows
    result.stackNeed = need;
    for (int a = 0; a < 3; ++a)
    {
        result.rootMin[a] = b.nodes[0].mn[a];
        result.rootMax[a] = b.nodes[0].mx[a];
    }
    *out = std::move(result);
    return true;
}

bool bvh4_adopt_device_tree(const float *aabbMin, const float *aabbMax, uint32_t count, const DeviceTree4 &tree,
                            Bvh4 *out)
{
    if (count < 2 || tree.slotPrim.size() != count || tree.nodes.empty() || tree.nodes.size() % 32 != 0) return false;
    const uint32_t nodeCount = (uint32_t)(tree.nodes.size() / 32);
    if (nodeCount > count - 1) return false;
    {
        std::vector<uint8_t> seen(count, 0);
        for (uint32_t p : tree.slotPrim)
        {
            if (p >= count || seen[p]) return false;
            seen[p] = 1;
        }
    }
    Bvh4 result;
    result.nodes.resize(nodeCount);
    static_assert(sizeof(Node4) == 32 * sizeof(uint32_t), "DeviceTree4 nodes are Node4s");
    memcpy(result.nodes.data(), tree.nodes.data(), (size_t)nodeCount * sizeof(Node4));
    result.slotPrim = tree.slotPrim;
    // what a node's parent says about it: the box it holds for it, its depth, the stack entries in use above it
    struct Entry { float mn[3], mx[3]; uint32_t depth, stackBefore; uint8_t referenced; };
    std::vector<Entry> entry(nodeCount);
    memset(entry.data(), 0, entry.size() * sizeof(Entry));
    for (int a = 0; a < 3; ++a) { entry[0].mn[a] = tree.rootBox[a]; entry[0].mx[a] = tree.rootBox[3 + a]; }
    entry[0].referenced = 1;
    std::vector<uint8_t> slotSeen(count, 0);
    uint32_t worstStack = 0, maxDepth = 0;
    for (uint32_t i = 0; i < nodeCount; ++i)
    {
        Node4 &nd = result.nodes[i];
        const Entry &me = entry[i];
        if (!me.referenced) return false; // (parents come before their children: an unreferenced node here stays so)
        const uint32_t n = nd.meta[0];
        if (n < 1 || n > 4 || nd.meta[1] != me.depth) return false;
        const uint32_t stackHere = me.stackBefore + (n - 1);
        worstStack = std::max(worstStack, stackHere);
        maxDepth = std::max(maxDepth, me.depth);
        for (uint32_t k = 0; k < 4; ++k)
        {
            const uint32_t ref = nd.ref[k];
            if (k >= n)
            {
                if (ref != SPB_REF_EMPTY) return false;
                for (int a = 0; a < 3; ++a) { nd.bmin[a][k] = NAN; nd.bmax[a][k] = NAN; }
                continue;
            }
            if (ref == SPB_REF_EMPTY) return false;
            if (ref & SPB_REF_LEAF)
            {
                const uint32_t slot = ref & ~SPB_REF_LEAF;
                if (slot >= count || slotSeen[slot]) return false;
                slotSeen[slot] = 1;
                const uint32_t prim = result.slotPrim[slot];
                for (int a = 0; a < 3; ++a) { nd.bmin[a][k] = aabbMin[(size_t)prim * 3 + a]; nd.bmax[a][k] = aabbMax[(size_t)prim * 3 + a]; }
            }
            else
            {
                if (ref <= i || ref >= nodeCount || entry[ref].referenced) return false;
                Entry &e = entry[ref];
                e.referenced = 1;
                e.depth = me.depth + 1;
                e.stackBefore = stackHere;
                for (int a = 0; a < 3; ++a) { e.mn[a] = nd.bmin[a][k]; e.mx[a] = nd.bmax[a][k]; }
            }
            // inside the box the parent holds for this node (NaN boxes aside, as in bvh4_from_binary)
            for (int a = 0; a < 3; ++a)
                if (nd.bmin[a][k] < me.mn[a] || nd.bmax[a][k] > me.mx[a]) return false;
        }
    }
    for (uint32_t slot = 0; slot < count; ++slot)
        if (!slotSeen[slot]) return false;
    if (worstStack > SPB_MESH_STACK_LIMIT) return false;
    result.maxDepth = maxDepth;
    result.stackNeed = worstStack;
    for (int a = 0; a < 3; ++a) { result.rootMin[a] = tree.rootBox[a]; result.rootMax[a] = tree.rootBox[3 + a]; }
    *out = std::move(result);
    return true;
}

bool lbvh_collapse_host_emulation(const float *aabbMin, const float *aabbMax, uint32_t count, const BinaryTree &tree,
                                  DeviceTree4 *out)
{
    if (count < 2) return false;
    const uint32_t internal = count - 1;
    if (tree.sortedPrim.size() != count || tree.children.size() != (size_t)internal * 2 || tree.boxes.size() != (size_t)internal * 6)
        return false;
    // the dynamic programme bottom-up, the way k_lbvh_fit reaches the nodes: a node after both its children
    std::vector<uint32_t> parent(internal, 0xFFFFFFFFu), pendingKids(internal, 0), ready;
    for (uint32_t i = 0; i < internal; ++i)
        for (int c = 0; c < 2; ++c)
        {
            uint32_t ref = tree.children[(size_t)i * 2 + c];
            if (ref & SPB_REF_LEAF) continue;
            if (ref >= internal) return false;
            parent[ref] = i;
            pendingKids[i]++;
        }
    for (uint32_t i = 0; i < internal; ++i)
        if (!pendingKids[i]) ready.push_back(i);
    std::vector<float> cost((size_t)internal * 3, 0.0f);
    std::vector<uint32_t> picks(internal, 0);
    size_t done = 0;
    while (done < ready.size())
    {
        uint32_t n = ready[done++];
        const uint32_t l = tree.children[(size_t)n * 2], r = tree.children[(size_t)n * 2 + 1];
        float zero[3] = {0.0f, 0.0f, 0.0f};
        lbvh_dp(&tree.boxes[(size_t)n * 6], (l & SPB_REF_LEAF) ? zero : &cost[(size_t)l * 3], (r & SPB_REF_LEAF) ? zero : &cost[(size_t)r * 3],
                &cost[(size_t)n * 3], &picks[n]);
        if (parent[n] != 0xFFFFFFFFu && --pendingKids[parent[n]] == 0) ready.push_back(parent[n]);
    }
    if (done != internal) return false;
    LbvhDpView view = {tree.children.data(), cost.data(), picks.data()};
    std::vector<uint32_t> nodes4((size_t)internal * 32, 0);
    out->slotPrim.assign(count, 0);
    uint32_t counters[5] = {1, 0, 0, 0, 0};
    std::vector<LbvhEmitItem> items[2];
    items[0].push_back({0, 0, 0, 0});
    for (uint32_t level = 0; !items[level & 1].empty(); ++level)
    {
        if (level >= 192) return false;
        std::vector<LbvhEmitItem> &cur = items[level & 1], &next = items[(level + 1) & 1];
        next.assign(internal, LbvhEmitItem());
        uint32_t nextCount = 0;
        for (const LbvhEmitItem &it : cur)
            lbvh_emit(view, tree.sortedPrim.data(), aabbMin, aabbMax, tree.boxes.data(), it, nodes4.data(), out->slotPrim.data(), counters,
                      next.data(), &nextCount);
        next.resize(nextCount);
    }
    if (counters[0] > internal || counters[1] != count) return false;
    nodes4.resize((size_t)counters[0] * 32);
    out->nodes = std::move(nodes4);
    out->maxDepth = counters[3];
    out->stackNeed = counters[4];
    for (int a = 0; a < 6; ++a) out->rootBox[a] = tree.boxes[a];
    return true;
}

Bvh4 build_bvh4_lbvh_host_device_collapse(const float *aabbMin, const float *aabbMax, uint32_t count)
{
    Bvh4 out;
    DeviceTree4 tree;
    if (count >= 2 && lbvh_collapse_host_emulation(aabbMin, aabbMax, count, lbvh_build_binary_host(aabbMin, aabbMax, count), &tree) &&
        bvh4_adopt_device_tree(aabbMin, aabbMax, count, tree, &out))
        return out;
    return build_bvh4(aabbMin, aabbMax, count);
}

Bvh4 build_bvh4_lbvh_host(const float *aabbMin, const float *aabbMax, uint32_t count)
{
    Bvh4 out;
    if (count >= 2 && bvh4_from_binary(aabbMin, aabbMax, count, lbvh_build_binary_host(aabbMin, aabbMax, count), &out))
        return out;
    return build_bvh4(aabbMin, aabbMax, count);
}

float bvh4_extent(const Bvh4 &bvh)
{
    float big = 0.0f;
    if (bvh.nodes.empty()) return big;
    const Node4 &n = bvh.nodes[0];
    for (int a = 0; a < 3; ++a)
        for (int k = 0; k < 4; ++k)
        {
            // NaN: an empty slot or a leaf with a NaN vertex (kept out of its ancestors); an infinite
            // corner makes the extent infinite, which sends every ray of the tree to the exact walk
            float lo = std::fabs(n.bmin[a][k]), hi = std::fabs(n.bmax[a][k]);
            if (lo > big) big = lo;
            if (hi > big) big = hi;
        }
    return big;
}

Bvh4 build_bvh4(const float *aabbMin, const float *aabbMax, uint32_t count, uint32_t stackLimit)
{
    Bvh4 out;
    if (count == 0) return out;
    if (stackLimit == 0) stackLimit = SPB_MESH_STACK_LIMIT;
    Builder b;
    b.aabbMin = aabbMin;
    b.aabbMax = aabbMax;
    b.balancedOnly = false;
    b.build(count);
    uint32_t need = collapse(b, out);
    // The traversal stack gives a mesh tree SPB_MESH_STACK_LIMIT entries and the TLAS
    // SPB_TLAS_STACK_LIMIT (spb_core.cuh).  A SAH tree that could exceed its share (pathological
    // inputs: e.g. thousands of objects along a line) is rebuilt with median splits, whose depth
    // is ceil(log2 n).
    if (need > stackLimit)
    {
        b.balancedOnly = true;
        b.build(count);
        need = collapse(b, out);
    }
    out.stackNeed = need;
    for (int a = 0; a < 3; ++a)
    {
        out.rootMin[a] = b.nodes[0].mn[a];
        out.rootMax[a] = b.nodes[0].mx[a];
    }
    return out;
}

} // namespace spb
