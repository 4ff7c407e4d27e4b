// spb_lbvh.cuh -- per-element arithmetic of the device BVH builder ("scalable
// BVH construction on device (LBVH -> 4-wide)", replacing bvh_CreateTree, src/bvh.cpp:51-200, whose
// agglomerative build is O(n^2 log n): 4.3 s for the 15 744-triangle monkey).
//
// Linear BVH after Karras, "Maximizing parallelism in the construction of BVHs, octrees, and k-d
// trees" (HPG 2012): 63-bit Morton codes of the primitive centroids, one radix sort, then every
// internal node of the binary radix tree finds its own range and split independently.  One
// definition for host and device: spb_lbvh.cu runs these per thread, spb_bvh.cpp runs them in a loop
// (the host emulation tests/hostsim checks without a GPU).
//
// Parity does not depend on the tree (DESIGN.md "Closest-hit definition"): every primitive ends up
// alone in a child slot with exactly its own AABB, whatever sits above it.
#pragma once
#include "spb_core.cuh"

namespace spb {

// spread the low 21 bits of v so that there are two zero bits between consecutive bits
SPB_HD uint64_t lbvh_expand21(uint64_t v)
{
    v &= 0x1FFFFFull;
    v = (v | (v << 32)) & 0x001F00000000FFFFull;
    v = (v | (v << 16)) & 0x001F0000FF0000FFull;
    v = (v | (v << 8)) & 0x100F00F00F00F00Full;
    v = (v | (v << 4)) & 0x10C30C30C30C30C3ull;
    v = (v | (v << 2)) & 0x1249249249249249ull;
    return v;
}

// 63-bit Morton code of the box centre, quantised to 2^21 cells per axis inside [rootMin, rootMax].
// A centre that is not finite, or an axis without extent, lands in cell 0 of that axis.
SPB_HD uint64_t lbvh_key(const float *mn, const float *mx, const float *rootMin, const float *rootMax)
{
    uint64_t cell[3];
    for (int a = 0; a < 3; ++a)
    {
        float extent = rootMax[a] - rootMin[a];
        float c = (mn[a] + mx[a]) * 0.5f;
        float u = extent > 0.0f ? (c - rootMin[a]) / extent : 0.0f;
        if (!(u >= 0.0f)) u = 0.0f; // NaN and negatives
        if (u > 1.0f) u = 1.0f;
        float q = u * 2097151.0f;
        cell[a] = (uint64_t)q;
        if (cell[a] > 2097151ull) cell[a] = 2097151ull;
    }
    return (lbvh_expand21(cell[0]) << 2) | (lbvh_expand21(cell[1]) << 1) | lbvh_expand21(cell[2]);
}

SPB_HD int lbvh_clz64(uint64_t v)
{
#if defined(__CUDA_ARCH__)
    return __clzll((long long)v);
#else
    return v ? __builtin_clzll(v) : 64;
#endif
}

// length of the common prefix of the keys at sorted positions i and j; equal keys are told apart
// by the positions themselves; -1 outside [0, n)
SPB_HD int lbvh_delta(const uint64_t *keys, int64_t n, int64_t i, int64_t j)
{
    if (j < 0 || j >= n) return -1;
    uint64_t a = keys[i], b = keys[j];
    if (a != b) return lbvh_clz64(a ^ b);
    return 64 + lbvh_clz64((uint64_t)i ^ (uint64_t)j);
}

// Children of internal node i (0 <= i < n - 1) of the binary radix tree over n sorted keys.
// A child is an internal node index, or SPB_REF_LEAF | sorted position for a leaf.
SPB_HD void lbvh_node(const uint64_t *keys, int64_t n, int64_t i, uint32_t &left, uint32_t &right)
{
    const int64_t d = lbvh_delta(keys, n, i, i + 1) - lbvh_delta(keys, n, i, i - 1) >= 0 ? 1 : -1;
    const int deltaMin = lbvh_delta(keys, n, i, i - d);
    int64_t lmax = 2;
    while (lbvh_delta(keys, n, i, i + lmax * d) > deltaMin) lmax *= 2;
    int64_t l = 0;
    for (int64_t t = lmax / 2; t >= 1; t /= 2)
        if (lbvh_delta(keys, n, i, i + (l + t) * d) > deltaMin) l += t;
    const int64_t j = i + l * d;
    const int deltaNode = lbvh_delta(keys, n, i, j);
    int64_t s = 0;
    for (int64_t t = (l + 1) / 2;; t = (t + 1) / 2)
    {
        if (lbvh_delta(keys, n, i, i + (s + t) * d) > deltaNode) s += t;
        if (t <= 1) break;
    }
    const int64_t gamma = i + s * d + (d < 0 ? -1 : 0);
    const int64_t lo = i < j ? i : j, hi = i < j ? j : i;
    left = lo == gamma ? (SPB_REF_LEAF | (uint32_t)gamma) : (uint32_t)gamma;
    right = hi == gamma + 1 ? (SPB_REF_LEAF | (uint32_t)(gamma + 1)) : (uint32_t)(gamma + 1);
}

// ------------------------------------------------------------------------------------------
// 4-wide collapse of the binary tree, per element (device form of spb_bvh.cpp WideDp / collapse: the
// dynamic programme of Ylitie, Karras and Laine, HPG 2017, section 4.1, for width 4 and single-primitive
// leaves).  Same arithmetic in the same order as the host collapse, so the two choose the same children
// for every 4-wide node; only the numbering of nodes and leaf slots inside a level differs on the GPU
// (atomic counters), which no result depends on.
//   cost[3 n + i - 1] = C(n, i), i = 1..3; picks[n] holds k of the best split for j = 2, 3, 4 roots, 2 bits each.
// A leaf reference (SPB_REF_LEAF | sorted position) costs 0 whatever i.

SPB_HD float lbvh_half_area(const float *b)
{
    float dx = b[3] - b[0], dy = b[4] - b[1], dz = b[5] - b[2];
    return dx * dy + dy * dz + dz * dx;
}

// C(n, .) and the picks of internal node n from its box and its children's costs (cl, cr: three floats
// each, zeros for a leaf child)
SPB_HD void lbvh_dp(const float *box, const float *cl, const float *cr, float *cost3, uint32_t *picks)
{
    const float inf = u2f(0x7F800000u);
    float dist[5] = {0, 0, 0, 0, 0};
    uint32_t p = 0;
    for (int j = 2; j <= 4; ++j)
    {
        float best = inf;
        uint32_t bestK = 1;
        for (int k = 1; k < j; ++k)
        {
            if (k > 3 || j - k > 3) continue;
            float c = cl[k - 1] + cr[j - k - 1];
            if (c < best) { best = c; bestK = (uint32_t)k; }
        }
        dist[j] = best;
        p |= bestK << (2 * (j - 2));
    }
    float a = lbvh_half_area(box);
    if (!(a >= 0.0f)) a = 0.0f;
    cost3[0] = a + dist[4];
    cost3[1] = cost3[0] < dist[2] ? cost3[0] : dist[2];
    cost3[2] = cost3[1] < dist[3] ? cost3[1] : dist[3];
    *picks = p;
}

struct LbvhDpView
{
    const uint32_t *children; // 2 per internal node
    const float *cost;        // 3 per internal node
    const uint32_t *picks;    // 1 per internal node
};
SPB_HD float lbvh_cost_of(const LbvhDpView &t, uint32_t ref, int i)
{
    return (ref & SPB_REF_LEAF) ? 0.0f : t.cost[(size_t)ref * 3 + (i - 1)];
}
SPB_HD int lbvh_pick_of(const LbvhDpView &t, uint32_t node, int j) { return (int)((t.picks[node] >> (2 * (j - 2))) & 3u); }

// The roots that stand for subtree `ref` when it may use up to j child slots, left to right.
SPB_HD void lbvh_gather(const LbvhDpView &t, uint32_t ref, int j, uint32_t *kids, uint32_t &count)
{
    uint32_t pendingRef[4];
    int pendingJ[4];
    int sp = 0;
    pendingRef[sp] = ref; pendingJ[sp] = j; sp++;
    while (sp)
    {
        sp--;
        uint32_t n = pendingRef[sp];
        int jj = pendingJ[sp];
        for (;;)
        {
            if ((n & SPB_REF_LEAF) || jj <= 1) { kids[count++] = n; break; }
            const uint32_t l = t.children[(size_t)n * 2], r = t.children[(size_t)n * 2 + 1];
            const int k = lbvh_pick_of(t, n, jj);
            if (jj <= 3)
            {
                // C(n, jj) = min(dist[jj], C(n, jj - 1)): fewer roots on ties
                float distJ = lbvh_cost_of(t, l, k) + lbvh_cost_of(t, r, jj - k);
                if (!(distJ < t.cost[(size_t)n * 3 + (jj - 2)])) { jj--; continue; }
            }
            pendingRef[sp] = r; pendingJ[sp] = jj - k; sp++;
            n = l;
            jj = k;
        }
    }
}

SPB_HD uint32_t lbvh_take(uint32_t *counter, uint32_t n)
{
#if defined(__CUDA_ARCH__)
    return atomicAdd(counter, n);
#else
    uint32_t v = *counter;
    *counter = v + n;
    return v;
#endif
}
SPB_HD void lbvh_raise(uint32_t *counter, uint32_t v)
{
#if defined(__CUDA_ARCH__)
    atomicMax(counter, v);
#else
    if (v > *counter) *counter = v;
#endif
}

// One 4-wide node of the result: `item` = (binary internal node that roots it, its index in the output, its depth,
// traversal stack entries in use above it).  Writes the node (layout of spb_bvh.h Node4: bmin[3][4], bmax[3][4],
// ref[4], meta[4]), numbers its internal children and its leaf slots from the counters and appends the
// children's items to the next level's list (`next`, counted in *nextCount).  counters: [0] nodes, [1] leaf slots,
// [2] unused, [3] deepest node, [4] worst-case stack entries.
struct LbvhEmitItem { uint32_t bnode, node4, depth, stackBefore; };
SPB_HD void lbvh_emit(const LbvhDpView &t, const uint32_t *sortedPrim, const float *primMin, const float *primMax,
                      const float *boxes, LbvhEmitItem item, uint32_t *nodes4, uint32_t *slotPrim, uint32_t *counters,
                      LbvhEmitItem *next, uint32_t *nextCount)
{
    uint32_t kids[4];
    uint32_t n = 0;
    {
        const uint32_t l = t.children[(size_t)item.bnode * 2], r = t.children[(size_t)item.bnode * 2 + 1];
        const int k = lbvh_pick_of(t, item.bnode, 4);
        lbvh_gather(t, l, k, kids, n);
        lbvh_gather(t, r, 4 - k, kids, n);
    }
    uint32_t leaves = 0, inner = 0;
    for (uint32_t k = 0; k < n; ++k)
        if (kids[k] & SPB_REF_LEAF) leaves++; else inner++;
    uint32_t slot = leaves ? lbvh_take(&counters[1], leaves) : 0u;
    uint32_t child4 = inner ? lbvh_take(&counters[0], inner) : 0u;
    uint32_t at = inner ? lbvh_take(nextCount, inner) : 0u;
    const uint32_t stackHere = item.stackBefore + (n > 0 ? n - 1 : 0);
    uint32_t *out = nodes4 + (size_t)item.node4 * 32;
    for (uint32_t k = 0; k < 4; ++k)
    {
        float mn[3], mx[3];
        uint32_t ref = SPB_REF_EMPTY;
        if (k >= n)
        {
            // empty lane: a NaN box never passes either form of the slab test
            for (int a = 0; a < 3; ++a) { mn[a] = u2f(0x7FC00000u); mx[a] = u2f(0x7FC00000u); }
        }
        else if (kids[k] & SPB_REF_LEAF)
        {
            const uint32_t prim = sortedPrim[kids[k] & ~SPB_REF_LEAF];
            for (int a = 0; a < 3; ++a) { mn[a] = primMin[(size_t)prim * 3 + a]; mx[a] = primMax[(size_t)prim * 3 + a]; }
            slotPrim[slot] = prim;
            ref = SPB_REF_LEAF | slot;
            slot++;
        }
        else
        {
            for (int a = 0; a < 3; ++a) { mn[a] = boxes[(size_t)kids[k] * 6 + a]; mx[a] = boxes[(size_t)kids[k] * 6 + 3 + a]; }
            ref = child4;
            LbvhEmitItem c;
            c.bnode = kids[k]; c.node4 = child4; c.depth = item.depth + 1; c.stackBefore = stackHere;
            next[at] = c;
            child4++;
            at++;
        }
        for (int a = 0; a < 3; ++a) { out[a * 4 + k] = f2u(mn[a]); out[12 + a * 4 + k] = f2u(mx[a]); }
        out[24 + k] = ref;
    }
    out[28] = n;
    out[29] = item.depth;
    out[30] = 0;
    out[31] = 0;
    lbvh_raise(&counters[3], item.depth);
    lbvh_raise(&counters[4], stackHere);
}

} // namespace spb
