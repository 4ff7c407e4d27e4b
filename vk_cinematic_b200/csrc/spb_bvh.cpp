// spb_bvh.cpp -- binned-SAH binary build, then a greedy collapse to 4 children per node.
// See spb_bvh.h for what is and is not kept from the reference's bvh_CreateTree.
#include "spb_bvh.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <queue>

#include "spb_core.cuh"

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

    // Returns the split position (first index of the right half) after partitioning `order`.
    uint32_t split(uint32_t first, uint32_t count, const float *nodeMin, const float *nodeMax)
    {
        const int BINS = 16;
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

// Greedy collapse in breadth-first order.  Returns the worst-case traversal stack depth.
uint32_t collapse(const Builder &b, Bvh4 &out)
{
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

Bvh4 build_bvh4(const float *aabbMin, const float *aabbMax, uint32_t count)
{
    Bvh4 out;
    if (count == 0) return out;
    Builder b;
    b.aabbMin = aabbMin;
    b.aabbMax = aabbMax;
    b.balancedOnly = false;
    b.build(count);
    uint32_t need = collapse(b, out);
    // The traversal stack gives each tree 2/3 of SPB_STACK_SIZE entries.  A SAH tree that could
    // exceed it (pathological inputs only) is rebuilt with median splits, whose depth is
    // ceil(log2 n).
    if (need + 4 > (SPB_STACK_SIZE * 2) / 3)
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
