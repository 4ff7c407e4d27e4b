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

} // namespace spb
