// oracle/sp_oracle.cpp -- CPU restatement ("port") of the reference's sp_ path.
//
// TEST INFRASTRUCTURE ONLY (see oracle/ora_api.h): loaded by tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs, never by the product.
//
// This is a scalar, single-source restatement of the algorithm of deadVertex/vk_cinematic's CPU
// path tracer, written against the reference sources (cited per function as file:line relative
// to the reference checkout) but sharing no code with them or with the CUDA library.  It exists
// for what the verbatim reference (oracle/_ref) cannot do: more than 3 bounces (literal at
// simd_path_tracer.cpp:195, path[4] at :233), more than 32 objects (sp_scene.h:15), and running
// on a machine where /root/reference is not mounted.
//
// Pinned: tests/test_oracle_port.py checks this file against (a) every known-answer value of the
// reference's own unit tests (tests/golden/reference_kats.json, transcribed with file:line),
// (b) golden images / hit maps produced by oracle/_ref here (tests/golden/*.npz, made by
// tools/make_golden.py), and (c) oracle/_ref itself, bit for bit, wherever that library is
// present.  Parity status: PINNED.
//
// Floating point: build with -O2 -ffp-contract=off and no -march flags; every expression keeps
// the reference's association order.  libm: sinf/cosf/atan2f/powf/sqrtf/floorf of the host
// glibc, exactly as the reference calls them (math_utils.h:45-121).  With
// -DORA_DETERMINISTIC_MATH the first four are replaced by "evaluate in double, round once"
// (same switch as oracle/ref_driver.cpp).
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <thread>
#include <vector>

#include "ora_api.h"

#ifdef ORA_DETERMINISTIC_MATH
static inline float lm_sin(float x) { return (float)sin((double)x); }
static inline float lm_cos(float x) { return (float)cos((double)x); }
static inline float lm_atan2(float y, float x) { return (float)atan2((double)y, (double)x); }
static inline float lm_pow(float x, float y)
{
    if (y == 5.0f)
    {
        double d = (double)x, d2 = d * d;
        return (float)(d2 * d2 * d);
    }
    return (float)pow((double)x, (double)y);
}
#define PORT_NAME "port-dm"
#else
static inline float lm_sin(float x) { return sinf(x); }
static inline float lm_cos(float x) { return cosf(x); }
static inline float lm_atan2(float y, float x) { return atan2f(y, x); }
static inline float lm_pow(float x, float y) { return powf(x, y); }
#define PORT_NAME "port"
#endif

#if defined(__x86_64__)
#include <x86intrin.h>
static inline uint64_t tick(void) { return __rdtsc(); }
#else
static inline uint64_t tick(void) { return 0; }
#endif

// math_utils.h:6-7
static const float kPi = 3.14159265359f;
static const float kEps = 1.1920928955078125e-07f; // FLT_EPSILON

// ---------------------------------------------------------------------------------------------
// scalar helpers with the reference's semantics
struct V2 { float x, y; };
struct V3 { float x, y, z; };
struct V4 { float x, y, z, w; };
struct M44 { V4 col[4]; };

static inline float lo2(float a, float b) { return a < b ? a : b; }  // Min, math_utils.h:9-13
static inline float hi2(float a, float b) { return a > b ? a : b; }  // Max, math_utils.h:15-19
static inline V3 P3(float x, float y, float z) { V3 v = {x, y, z}; return v; }
static inline V3 plus(V3 a, V3 b) { return P3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline V3 minus(V3 a, V3 b) { return P3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline V3 times(V3 a, float s) { return P3(a.x * s, a.y * s, a.z * s); }
static inline V3 flip(V3 a) { return P3(-a.x, -a.y, -a.z); }
static inline V3 mulc(V3 a, V3 b) { return P3(a.x * b.x, a.y * b.y, a.z * b.z); } // Hadamard
static inline float inner(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline V3 outer(V3 a, V3 b) // Cross, math_lib.h:467-477
{
    return P3((a.y * b.z) - (a.z * b.y), (a.z * b.x) - (a.x * b.z), (a.x * b.y) - (a.y * b.x));
}
static inline V3 lo3(V3 a, V3 b) { return P3(lo2(a.x, b.x), lo2(a.y, b.y), lo2(a.z, b.z)); }
static inline V3 hi3(V3 a, V3 b) { return P3(hi2(a.x, b.x), hi2(a.y, b.y), hi2(a.z, b.z)); }
static inline V3 unit(V3 v) // Normalize, math_lib.h:503-512
{
    float len = sqrtf(inner(v, v));
    V3 r = {0, 0, 0};
    if (len > kEps) r = times(v, 1.0f / len);
    return r;
}
static inline float get4(const V4 &v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }
static inline void put4(V4 &v, int i, float f) { if (i == 0) v.x = f; else if (i == 1) v.y = f; else if (i == 2) v.z = f; else v.w = f; }
static inline float inner4(V4 a, V4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }

// math_lib.h:381-398: result[i] = Dot(row i, b)
static V4 apply(const M44 &m, V4 b)
{
    V4 r = {0, 0, 0, 0};
    for (int i = 0; i < 4; ++i)
    {
        V4 row = {get4(m.col[0], i), get4(m.col[1], i), get4(m.col[2], i), get4(m.col[3], i)};
        put4(r, i, inner4(row, b));
    }
    return r;
}
static inline V3 apply_point(V3 p, const M44 &m) { V4 r = apply(m, V4{p.x, p.y, p.z, 1.0f}); return P3(r.x, r.y, r.z); }
static inline V3 apply_vector(V3 d, const M44 &m) { V4 r = apply(m, V4{d.x, d.y, d.z, 0.0f}); return P3(r.x, r.y, r.z); }
// math_lib.h:250-272
static M44 compose(const M44 &a, const M44 &b)
{
    M44 r;
    memset(&r, 0, sizeof(r));
    for (int i = 0; i < 4; ++i)
    {
        V4 row = {get4(a.col[0], i), get4(a.col[1], i), get4(a.col[2], i), get4(a.col[3], i)};
        for (int j = 0; j < 4; ++j) put4(r.col[j], i, inner4(row, b.col[j]));
    }
    return r;
}
static M44 m_identity()
{
    M44 m;
    memset(&m, 0, sizeof(m));
    m.col[0].x = m.col[1].y = m.col[2].z = m.col[3].w = 1.0f;
    return m;
}
static M44 m_scale(V3 s) { M44 m; memset(&m, 0, sizeof(m)); m.col[0].x = s.x; m.col[1].y = s.y; m.col[2].z = s.z; m.col[3].w = 1.0f; return m; }
static M44 m_translate(V3 t) { M44 m = m_identity(); m.col[3].x = t.x; m.col[3].y = t.y; m.col[3].z = t.z; return m; }
// Rotate(quat), math_lib.h:623-650
static M44 m_rotate(V4 q)
{
    M44 m = m_identity();
    float x = q.x, y = q.y, z = q.z, w = q.w;
    m.col[0].x = 1.0f - 2.0f * y * y - 2.0f * z * z;
    m.col[0].y = 2.0f * x * y + 2.0f * z * w;
    m.col[0].z = 2.0f * x * z - 2.0f * y * w;
    m.col[1].x = 2.0f * x * y - 2.0f * z * w;
    m.col[1].y = 1.0f - 2.0f * x * x - 2.0f * z * z;
    m.col[1].z = 2.0f * y * z + 2.0f * x * w;
    m.col[2].x = 2.0f * x * z + 2.0f * y * w;
    m.col[2].y = 2.0f * y * z - 2.0f * x * w;
    m.col[2].z = 1.0f - 2.0f * x * x - 2.0f * y * y;
    return m;
}
static inline V4 q_conj(V4 q) { return V4{-q.x, -q.y, -q.z, q.w}; }
// quat product, math_lib.h:580-586
static V4 q_mul(V4 p, V4 q)
{
    V3 pv = P3(p.x, p.y, p.z), qv = P3(q.x, q.y, q.z);
    V3 v = plus(plus(times(qv, p.w), times(pv, q.w)), outer(pv, qv));
    return V4{v.x, v.y, v.z, p.w * q.w - inner(pv, qv)};
}
// RotateVector, math_lib.h:608-613
static V3 q_rotate(V3 v, V4 p)
{
    V4 r = q_mul(q_mul(p, V4{v.x, v.y, v.z, 0.0f}), q_conj(p));
    return P3(r.x, r.y, r.z);
}

// ---------------------------------------------------------------------------------------------
// RNG, math_utils.h:184-214
static inline uint32_t xs32(uint32_t *s)
{
    uint32_t x = *s;
    x ^= x << 13;
    x ^= x >> 17;
    x ^= x << 5;
    *s = x;
    return x;
}
static inline float rnd01(uint32_t *s)
{
    float num = (float)(xs32(s) >> 1);
    float den = (float)(0xFFFFFFFFu >> 1);
    return num / den;
}
static inline float rnd11(uint32_t *s) { return -1.0f + 2.0f * rnd01(s); }

// ---------------------------------------------------------------------------------------------
// ray primitives

// simd_RayIntersectAabb4, simd.h:198-271 -- one lane.  minps/maxps return the second operand
// when the comparison is unordered, which `a < b ? a : b` reproduces.
static inline bool slab_lane(V3 bmin, V3 bmax, V3 o, V3 inv)
{
    float t0x = (bmin.x - o.x) * inv.x, t1x = (bmax.x - o.x) * inv.x;
    float t0y = (bmin.y - o.y) * inv.y, t1y = (bmax.y - o.y) * inv.y;
    float t0z = (bmin.z - o.z) * inv.z, t1z = (bmax.z - o.z) * inv.z;
    V3 tmin = P3(t0x < t1x ? t0x : t1x, t0y < t1y ? t0y : t1y, t0z < t1z ? t0z : t1z);
    V3 tmax = P3(t0x > t1x ? t0x : t1x, t0y > t1y ? t0y : t1y, t0z > t1z ? t0z : t1z);
    float enter = hi2(0.0f, hi2(tmin.x, hi2(tmin.y, tmin.z)));
    float leave = lo2(tmax.x, lo2(tmax.y, tmax.z));
    return enter <= leave;
}

// RayIntersectAabb, ray_intersection.cpp:24-77
static float slab_scalar(V3 bmin, V3 bmax, V3 o, V3 d)
{
    float tmin = 0.0f, tmax = 3.402823466e+38f;
    const float *mn = &bmin.x, *mx = &bmax.x, *os = &o.x, *ds = &d.x;
    for (int axis = 0; axis < 3; ++axis)
    {
        if (fabsf(ds[axis]) < kEps)
        {
            if (os[axis] < mn[axis] || os[axis] > mx[axis]) return -1.0f;
        }
        else
        {
            float a = (mn[axis] - os[axis]) / ds[axis];
            float b = (mx[axis] - os[axis]) / ds[axis];
            float t0 = lo2(a, b), t1 = hi2(a, b);
            if (t0 > tmin) tmin = t0;
            tmax = lo2(tmax, t1);
            if (tmin > tmax) return -1.0f;
        }
    }
    return tmin;
}

struct TriHit { float t; V2 uv; V3 n; };

// RayIntersectTriangleMT, ray_intersection.cpp:156-190
static TriHit moller_trumbore(V3 o, V3 d, V3 a, V3 b, V3 c)
{
    TriHit r;
    memset(&r, 0, sizeof(r));
    r.t = -1.0f;
    V3 T = minus(o, a), e1 = minus(b, a), e2 = minus(c, a);
    V3 p = outer(d, e2), q = outer(T, e1), n = outer(e1, e2);
    V3 m = P3(inner(q, e2), inner(p, T), inner(q, d));
    float det = 1.0f / inner(p, e1);
    float t = det * m.x, u = det * m.y, v = det * m.z;
    float w = 1.0f - u - v;
    float f = inner(d, n);
    if (u >= 0.0f && u <= 1.0f && v >= 0.0f && v <= 1.0f && w >= 0.0f && w <= 1.0f && f < 0.0f)
    {
        r.t = t;
        r.n = n;
        r.uv = V2{u, v};
    }
    return r;
}

// ---------------------------------------------------------------------------------------------
// bvh_CreateTree / bvh_IntersectRay, bvh.cpp:51-311 -- nodes by index instead of by pointer
struct TreeNode { V3 mn, mx; int child[4]; uint32_t leaf; };
struct Tree { std::vector<TreeNode> nodes; int root = -1; };

struct SortEntry { float distSq; int64_t node; }; // 16 bytes like bvh_NodeDistSqPair
static int by_distance_desc(const void *pa, const void *pb)
{
    const SortEntry *a = (const SortEntry *)pa, *b = (const SortEntry *)pb;
    if (a->distSq < b->distSq) return 1;
    if (a->distSq > b->distSq) return -1;
    return 0;
}

static Tree grow_tree(const V3 *mins, const V3 *maxs, uint32_t count)
{
    Tree tree;
    if (count == 0) return tree;
    tree.nodes.reserve((size_t)count * 2);
    std::vector<SortEntry> pending[2];
    pending[0].resize(count);
    pending[1].resize(count);
    uint32_t size[2] = {count, 0};
    int rd = 0, wr = 1;
    for (uint32_t i = 0; i < count; ++i)
    {
        TreeNode n;
        n.mn = mins[i];
        n.mx = maxs[i];
        n.child[0] = n.child[1] = n.child[2] = n.child[3] = -1;
        n.leaf = i;
        tree.nodes.push_back(n);
        pending[rd][i].node = (int64_t)i;
        pending[rd][i].distSq = 0.0f;
    }
    int last = (int)pending[rd][0].node;
    for (;;)
    {
        size[wr] = 0;
        while (size[rd] > 0)
        {
            uint32_t top = size[rd] - 1;
            int self = (int)pending[rd][top].node;
            V3 centre = times(plus(tree.nodes[self].mx, tree.nodes[self].mn), 0.5f);
            // sort everything below the top entry, farthest first (bvh.cpp:36-49,117-118)
            for (uint32_t i = 0; i < top; ++i)
            {
                const TreeNode &o = tree.nodes[(int)pending[rd][i].node];
                V3 c = times(plus(o.mx, o.mn), 0.5f);
                V3 d = minus(c, centre);
                pending[rd][i].distSq = inner(d, d);
            }
            qsort(pending[rd].data(), top, sizeof(SortEntry), by_distance_desc);

            uint32_t near = top < 3 ? top : 3;
            if (near > 0)
            {
                TreeNode parent;
                parent.leaf = 0xFFFFFFFFu;
                parent.mn = tree.nodes[self].mn;
                parent.mx = tree.nodes[self].mx;
                parent.child[0] = self;
                parent.child[1] = parent.child[2] = parent.child[3] = -1;
                for (uint32_t i = 0; i < near; ++i)
                {
                    int nb = (int)pending[rd][top - (i + 1)].node;
                    parent.mn = lo3(parent.mn, tree.nodes[nb].mn);
                    parent.mx = hi3(parent.mx, tree.nodes[nb].mx);
                    parent.child[i + 1] = nb;
                }
                last = (int)tree.nodes.size();
                tree.nodes.push_back(parent);
                pending[wr][size[wr]++].node = last;
                size[rd] -= near + 1;
            }
            else
            {
                pending[wr][size[wr]++].node = self;
                size[rd]--;
            }
        }
        if (size[wr] > 1) { int t = rd; rd = wr; wr = t; }
        else break;
    }
    tree.root = last;
    return tree;
}

// Scalable builder for inputs the reference's O(n^2 log n) agglomeration cannot finish (the
// C5-style scenes, which the reference cannot run at all: 32-object table, sp_scene.h:15).  The set
// of leaves a ray reports does not depend on the tree above them (a leaf is reported iff its box
// chain passes, and the chain passes whenever the leaf's own box does -- SURVEY.md 7.2), so a
// median-split 4-ary tree gives the same closest hit; only the order among exactly equal t can
// differ from what the reference's tree would give.
static int grow_fast_rec(Tree &tree, const V3 *mins, const V3 *maxs, std::vector<uint32_t> &order,
                         uint32_t first, uint32_t count)
{
    if (count == 1)
    {
        TreeNode n;
        n.mn = mins[order[first]];
        n.mx = maxs[order[first]];
        n.child[0] = n.child[1] = n.child[2] = n.child[3] = -1;
        n.leaf = order[first];
        tree.nodes.push_back(n);
        return (int)tree.nodes.size() - 1;
    }
    auto split = [&](uint32_t f, uint32_t c) -> uint32_t {
        V3 lo = P3(INFINITY, INFINITY, INFINITY), hi = P3(-INFINITY, -INFINITY, -INFINITY);
        for (uint32_t i = f; i < f + c; ++i)
        {
            V3 ctr = times(plus(mins[order[i]], maxs[order[i]]), 0.5f);
            lo = lo3(lo, ctr);
            hi = hi3(hi, ctr);
        }
        V3 e = minus(hi, lo);
        int axis = e.y > e.x ? (e.z > e.y ? 2 : 1) : (e.z > e.x ? 2 : 0);
        uint32_t mid = f + c / 2;
        std::nth_element(order.begin() + f, order.begin() + mid, order.begin() + f + c,
            [&](uint32_t a, uint32_t b) {
                float ca = axis == 0 ? mins[a].x + maxs[a].x : axis == 1 ? mins[a].y + maxs[a].y : mins[a].z + maxs[a].z;
                float cb = axis == 0 ? mins[b].x + maxs[b].x : axis == 1 ? mins[b].y + maxs[b].y : mins[b].z + maxs[b].z;
                return ca < cb || (ca == cb && a < b);
            });
        return mid;
    };
    uint32_t part[5] = {first, 0, 0, 0, first + count};
    part[2] = split(first, count);
    part[1] = part[2] - first >= 2 ? split(first, part[2] - first) : part[2];
    part[3] = part[4] - part[2] >= 2 ? split(part[2], part[4] - part[2]) : part[4];
    TreeNode parent;
    parent.leaf = 0xFFFFFFFFu;
    parent.child[0] = parent.child[1] = parent.child[2] = parent.child[3] = -1;
    parent.mn = P3(INFINITY, INFINITY, INFINITY);
    parent.mx = P3(-INFINITY, -INFINITY, -INFINITY);
    int kids = 0;
    for (int k = 0; k < 4; ++k)
    {
        if (part[k + 1] == part[k]) continue;
        int c = grow_fast_rec(tree, mins, maxs, order, part[k], part[k + 1] - part[k]);
        parent.child[kids++] = c;
        parent.mn = lo3(parent.mn, tree.nodes[c].mn);
        parent.mx = hi3(parent.mx, tree.nodes[c].mx);
    }
    tree.nodes.push_back(parent);
    return (int)tree.nodes.size() - 1;
}

static const uint32_t kFastBuildThreshold = 24000; // above the monkey (15 744 triangles)

static Tree grow_tree_any(const V3 *mins, const V3 *maxs, uint32_t count)
{
    if (count <= kFastBuildThreshold) return grow_tree(mins, maxs, count);
    Tree tree;
    tree.nodes.reserve((size_t)count * 2);
    std::vector<uint32_t> order(count);
    for (uint32_t i = 0; i < count; ++i) order[i] = i;
    tree.root = grow_fast_rec(tree, mins, maxs, order, 0, count);
    return tree;
}

struct Query { uint32_t count; uint32_t aabbTests; bool overflow; };

// Level-order walk with two ping-pong stacks; every leaf whose box chain passes is reported
// (bvh.cpp:203-311).  `cap` = 0 means unbounded (the port never aborts on the 128/32 caps).
static Query walk_tree(const Tree &tree, V3 o, V3 d, std::vector<int> &out, uint32_t cap,
                       std::vector<int> *scratch)
{
    Query q = {0, 0, false};
    out.clear();
    if (tree.root < 0) return q;
    std::vector<int> &a = scratch[0], &b = scratch[1];
    a.clear();
    b.clear();
    V3 inv = P3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z); // Inverse, math_lib.h:911-915
    const TreeNode &root = tree.nodes[tree.root];
    if (slab_lane(root.mn, root.mx, o, inv)) a.push_back(tree.root);
    std::vector<int> *rd = &a, *wr = &b;
    while (!rd->empty())
    {
        int ni = rd->back();
        rd->pop_back();
        const TreeNode &n = tree.nodes[ni];
        if (n.child[0] >= 0)
        {
            uint32_t kids = 0;
            for (int i = 0; i < 4; ++i) if (n.child[i] >= 0) kids++;
            q.aabbTests += kids;
            for (uint32_t i = 0; i < kids; ++i)
            {
                const TreeNode &c = tree.nodes[n.child[i]];
                if (slab_lane(c.mn, c.mx, o, inv)) wr->push_back(n.child[i]);
            }
        }
        else
        {
            if (cap == 0 || q.count < cap)
            {
                out.push_back(ni);
                q.count++;
            }
            else
            {
                q.overflow = true;
                break;
            }
        }
        if (rd->empty()) { std::vector<int> *t = rd; rd = wr; wr = t; }
    }
    return q;
}

// ---------------------------------------------------------------------------------------------
// scene

struct Vertex { V3 p, n; V2 uv; }; // VertexPNT, mesh.h:11-16

struct Mesh
{
    std::vector<Vertex> verts;
    std::vector<uint32_t> idx;
    Tree mid;
    bool smooth;
};

struct Object
{
    int mesh;
    uint32_t material;
    M44 model, invModel;
    V3 bmin, bmax;
};

struct Image { std::vector<float> px; uint32_t w, h; };
struct MaterialRec { V3 albedo; uint32_t albedoTex; V3 emission; uint32_t emissionTex; float roughness; };

struct Camera
{
    V3 right, up, forward, position, filmCenter;
    float halfPixelW, halfPixelH, halfFilmW, halfFilmH;
    uint32_t width, height;
};

struct ora_Scene
{
    std::vector<Mesh> meshes;
    std::vector<Object> objects;
    Tree broad;
    std::vector<uint32_t> matKeys;
    std::vector<MaterialRec> mats;
    std::vector<uint32_t> imgKeys;
    std::vector<Image> imgs;
    uint32_t background = 0;
    Camera cam;
};

struct Counters64 { uint64_t v[ORA_METRIC_COUNT]; };

struct Scratch
{
    std::vector<int> leaves, objLeaves, st[2], st2[2];
    // tie bookkeeping for ora_tie_mask(): how many scene queries ended with two candidates of
    // exactly equal t (the only case where visiting order picks the winner)
    uint64_t ties = 0;
    bool meshTie = false;
};

// TransformAabb, aabb.h:29-58
static void move_box(V3 bmin, V3 bmax, V3 pos, V4 rot, V3 scl, V3 *omin, V3 *omax)
{
    M44 m = compose(compose(m_translate(pos), m_rotate(rot)), m_scale(scl));
    V3 corner[8] = {P3(bmin.x, bmin.y, bmin.z), P3(bmax.x, bmin.y, bmin.z), P3(bmax.x, bmin.y, bmax.z),
                    P3(bmin.x, bmin.y, bmax.z), P3(bmin.x, bmax.y, bmin.z), P3(bmax.x, bmax.y, bmin.z),
                    P3(bmax.x, bmax.y, bmax.z), P3(bmin.x, bmax.y, bmax.z)};
    V3 lo = apply_point(corner[0], m), hi = lo;
    for (int i = 1; i < 8; ++i)
    {
        V3 p = apply_point(corner[i], m);
        lo = lo3(lo, p);
        hi = hi3(hi, p);
    }
    *omin = lo;
    *omax = hi;
}

// sp_ConfigureCamera, simd_path_tracer.cpp:1-36
static void setup_camera(Camera *c, V3 position, V4 rotation, float filmDistance, uint32_t w, uint32_t h)
{
    c->width = w;
    c->height = h;
    c->position = position;
    c->right = q_rotate(P3(1, 0, 0), rotation);
    c->up = q_rotate(P3(0, 1, 0), rotation);
    c->forward = q_rotate(P3(0, 0, -1), rotation);
    c->filmCenter = plus(position, times(c->forward, filmDistance));
    c->halfPixelW = 0.5f / (float)w;
    c->halfPixelH = 0.5f / (float)h;
    float fw = 1.0f, fh = 1.0f;
    if (w > h) fh = (float)h / (float)w;
    else if (w < h) fw = (float)w / (float)h;
    c->halfFilmW = 0.5f * fw;
    c->halfFilmH = 0.5f * fh;
}

// sp_CalculateFilmPositions, simd_path_tracer.cpp:38-63
static V3 film_point(const Camera &c, V2 pixel)
{
    float fx = pixel.x / (float)c.width;
    float fy = pixel.y / (float)c.height;
    fy = 1.0f - fy;
    fx = fx * 2.0f - 1.0f;
    fy = fy * 2.0f - 1.0f;
    V3 p = times(c.right, c.halfFilmW * fx);
    p = plus(p, times(c.up, c.halfFilmH * fy));
    p = plus(p, c.filmCenter);
    return p;
}

struct MeshHit { TriHit tri; int triangle; };

// sp_RayIntersectMesh, sp_scene.cpp:127-227
static MeshHit hit_mesh(const Mesh &mesh, V3 o, V3 d, Counters64 *m, Scratch *sc)
{
    MeshHit best;
    memset(&best, 0, sizeof(best));
    best.tri.t = -1.0f;
    best.triangle = -1;
    uint64_t t0 = tick();
    Query q = walk_tree(mesh.mid, o, d, sc->leaves, 0, sc->st);
    m->v[ORA_METRIC_CYC_MIDPHASE] += tick() - t0;
    m->v[ORA_METRIC_MIDPHASE_AABB_TESTS] += q.aabbTests;
    for (uint32_t i = 0; i < q.count; ++i)
    {
        uint32_t tri = mesh.mid.nodes[sc->leaves[i]].leaf;
        const Vertex &v0 = mesh.verts[mesh.idx[tri * 3 + 0]];
        const Vertex &v1 = mesh.verts[mesh.idx[tri * 3 + 1]];
        const Vertex &v2 = mesh.verts[mesh.idx[tri * 3 + 2]];
        uint64_t t1 = tick();
        TriHit h = moller_trumbore(o, d, v0.p, v1.p, v2.p);
        m->v[ORA_METRIC_CYC_TRIANGLE] += tick() - t1;
        if (h.t > 0.0f)
        {
            if (h.t == best.tri.t) sc->meshTie = true; // equal t: the first one visited stays
            if (h.t < best.tri.t || best.tri.t < 0.0f)
            {
                sc->meshTie = false;
                best.tri = h;
                best.triangle = (int)tri;
                float w = 1.0f - h.uv.x - h.uv.y;
                V2 uv;
                uv.x = v0.uv.x * w + v1.uv.x * h.uv.x + v2.uv.x * h.uv.y;
                uv.y = v0.uv.y * w + v1.uv.y * h.uv.x + v2.uv.y * h.uv.y;
                if (mesh.smooth)
                    best.tri.n = unit(plus(plus(times(v0.n, w), times(v1.n, h.uv.x)), times(v2.n, h.uv.y)));
                best.tri.uv = uv;
            }
        }
    }
    return best;
}

struct SceneHit { float t; uint32_t material; V3 n; V2 uv; int object, triangle; };

// sp_RayIntersectScene, sp_scene.cpp:229-339
static SceneHit hit_scene(const ora_Scene *s, V3 o, V3 d, Counters64 *m, Scratch *sc)
{
    uint64_t start = tick();
    SceneHit r;
    memset(&r, 0, sizeof(r));
    r.t = -1.0f;
    r.object = r.triangle = -1;
    uint64_t b0 = tick();
    bool tied = false;
    Query q = walk_tree(s->broad, o, d, sc->objLeaves, 0, sc->st2);
    m->v[ORA_METRIC_CYC_BROADPHASE] += tick() - b0;
    for (uint32_t i = 0; i < q.count; ++i)
    {
        uint32_t oi = s->broad.nodes[sc->objLeaves[i]].leaf;
        const Object &ob = s->objects[oi];
        V3 lo = apply_point(o, ob.invModel);
        V3 ld = unit(apply_vector(d, ob.invModel));
        uint64_t m0 = tick();
        sc->meshTie = false;
        MeshHit mh = hit_mesh(s->meshes[ob.mesh], lo, ld, m, sc);
        m->v[ORA_METRIC_CYC_MESH] += tick() - m0;
        m->v[ORA_METRIC_MESH_TESTS]++;
        if (mh.tri.t >= 0.0f)
        {
            V3 lh = plus(lo, times(ld, mh.tri.t));
            V3 wh = apply_point(lh, ob.model);
            float t = inner(minus(wh, o), d);
            V3 wn = unit(apply_vector(mh.tri.n, ob.model));
            if (t == r.t) tied = true;
            if (t < r.t || r.t < 0.0f)
            {
                tied = sc->meshTie;
                r.t = t;
                r.material = ob.material;
                r.n = wn;
                r.uv = mh.tri.uv;
                r.object = (int)oi;
                r.triangle = mh.triangle;
            }
        }
    }
    m->v[ORA_METRIC_CYC_SCENE] += tick() - start;
    if (tied) sc->ties++;
    return r;
}

// SampleImageNearest, image.h:3-18.  The reference indexes unclamped; an index past the last
// texel (uv.y == 1) is an out-of-bounds read there.  Both checkers and the GPU define it as the
// last texel (ref_driver.cpp pads the image it gives the reference accordingly).
static V4 nearest(const Image &img, V2 uv)
{
    float fx = uv.x * img.w, fy = uv.y * img.h;
    float flx = floorf(fx), fly = floorf(fy);
    uint32_t x = flx > 0.0f ? (flx < 4294967040.0f ? (uint32_t)flx : 0xFFFFFF00u) : 0u;
    uint32_t y = fly > 0.0f ? (fly < 4294967040.0f ? (uint32_t)fly : 0xFFFFFF00u) : 0u;
    uint64_t i = (uint64_t)y * img.w + x, last = (uint64_t)img.w * img.h - 1;
    if (i > last) i = last;
    const float *p = &img.px[i * 4];
    return V4{p[0], p[1], p[2], p[3]};
}

// SampleImageBilinear, image.h:34-73
static V4 bilinear(const Image &img, V2 uv)
{
    float px = (uv.x * img.w) - 0.5f, py = (uv.y * img.h) - 0.5f;
    px = hi2(px, 0.0f);
    py = hi2(py, 0.0f);
    uint32_t x0 = (uint32_t)floorf(px), x1 = x0 + 1;
    uint32_t y0 = (uint32_t)floorf(py), y1 = y0 + 1;
    float fx = px - (float)x0, fy = py - (float)y0;
    x1 = x1 < img.w - 1 ? x1 : img.w - 1;
    y1 = y1 < img.h - 1 ? y1 : img.h - 1;
    auto at = [&](uint32_t x, uint32_t y) { const float *p = &img.px[((size_t)y * img.w + x) * 4]; return V4{p[0], p[1], p[2], p[3]}; };
    auto mix = [](V4 a, V4 b, float t) { float s = 1.0f - t; return V4{a.x * s + b.x * t, a.y * s + b.y * t, a.z * s + b.z * t, a.w * s + b.w * t}; };
    V4 top = mix(at(x0, y0), at(x1, y0), fx);
    V4 bot = mix(at(x0, y1), at(x1, y1), fx);
    return mix(top, bot, fy);
}

// ToSphericalCoordinates / MapToEquirectangular / MapSphericalToCartesianCoordinates,
// math_lib.h:849-884
static V2 to_sphere(V3 v)
{
    float inc = lm_atan2(sqrtf(v.x * v.x + v.z * v.z), v.y);
    float az = lm_atan2(v.z, v.x);
    return V2{az, inc};
}
static V2 to_equirect(V2 sc)
{
    if (sc.x < 0.0f) sc.x += 2.0f * kPi;
    V2 uv;
    uv.x = sc.x / (2.0f * kPi);
    uv.y = lm_cos(sc.y) * 0.5f + 0.5f;
    return uv;
}
static V3 from_sphere(V2 sc)
{
    V3 r;
    r.x = lm_sin(sc.y) * lm_cos(sc.x);
    r.z = lm_sin(sc.y) * lm_sin(sc.x);
    r.y = lm_cos(sc.y);
    return r;
}

// RandomDirectionOnHemisphere, math_lib.h:932-946
static V3 hemisphere(V3 n, uint32_t *rng)
{
    float theta = kPi * rnd01(rng);
    float phi = kPi * rnd11(rng);
    V3 d = from_sphere(V2{phi, theta});
    if (inner(d, n) < 0.0f) d = flip(d);
    return d;
}

struct PathVertex { uint32_t material; V3 position, out, in, n; V2 uv; }; // sp_material_system.h:22-30

static const MaterialRec *find_material(const ora_Scene *s, uint32_t id)
{
    for (size_t i = 0; i < s->matKeys.size(); ++i) if (s->matKeys[i] == id) return &s->mats[i];
    return NULL;
}
static const Image *find_image(const ora_Scene *s, uint32_t id)
{
    for (size_t i = 0; i < s->imgKeys.size(); ++i) if (s->imgKeys[i] == id) return &s->imgs[i];
    return NULL;
}

struct Shade { V3 albedo, emission; float roughness; };

// sp_EvaluateMaterial, sp_material_system.cpp:59-105
static Shade shade_material(const ora_Scene *s, const MaterialRec *mat, const PathVertex *v)
{
    Shade out;
    memset(&out, 0, sizeof(out));
    const Image *at = find_image(s, mat->albedoTex), *et = find_image(s, mat->emissionTex);
    if (at) { V4 c = nearest(*at, v->uv); out.albedo = P3(c.x, c.y, c.z); }
    else out.albedo = mat->albedo;
    if (et)
    {
        V2 uv = to_equirect(to_sphere(flip(v->out)));
        uv.y = 1.0f - uv.y;
        V4 c = nearest(*et, uv);
        V3 rad = P3(c.x, c.y, c.z);
        rad = times(rad, 1.0f);
        out.emission = rad;
    }
    else out.emission = mat->emission;
    out.roughness = mat->roughness;
    return out;
}

// ComputeRadianceForPath with its BRDF helpers, simd_path_tracer.cpp:65-175
static V3 path_radiance(const ora_Scene *s, const PathVertex *path, uint32_t n, float clampTo)
{
    V3 radiance = {0, 0, 0};
    for (int i = (int)n - 1; i >= 0; --i)
    {
        const PathVertex *v = path + i;
        Shade sh;
        memset(&sh, 0, sizeof(sh));
        const MaterialRec *mat = find_material(s, v->material);
        if (mat) sh = shade_material(s, mat, v);
        else sh.emission = P3(1, 0, 1);

        float cosine = hi2(0.0f, inner(v->n, v->in));
        V3 incoming = radiance;
        if (clampTo > 0.0f)
        {
            incoming.x = lo2(hi2(incoming.x, 0.0f), clampTo);
            incoming.y = lo2(hi2(incoming.y, 0.0f), clampTo);
            incoming.z = lo2(hi2(incoming.z, 0.0f), clampTo);
        }
        V3 L = v->in, N = v->n, V = v->out;
        V3 H = unit(plus(L, V));
        // FresnelSchlick
        float pw = lm_pow(1.0f - hi2(inner(H, V), 0.0f), 5.0f);
        V3 F0 = P3(0.04f, 0.04f, 0.04f);
        V3 F = plus(F0, times(minus(P3(1, 1, 1), F0), pw));
        V3 kD = minus(P3(1, 1, 1), F);
        float invPi = 1.0f / kPi;
        float rough = sh.roughness;
        // DistributionGGX
        float a = rough * rough, a2 = a * a;
        float NdotH = hi2(inner(N, H), 0.0f), NdotH2 = NdotH * NdotH;
        float den = (NdotH2 * (a2 - 1.0f) + 1.0f);
        den = kPi * den * den;
        float NDF = a2 / den;
        // GeometrySmith
        float NdotV = hi2(inner(N, V), 0.0f), NdotL = hi2(inner(N, L), 0.0f);
        float rr = (rough + 1.0f), k = (rr * rr) / 8.0f;
        float g2 = NdotV / (NdotV * (1.0f - k) + k);
        float g1 = NdotL / (NdotL * (1.0f - k) + k);
        float G = g1 * g2;
        V3 num = times(F, NDF * G);
        float dn = 4.0f * hi2(inner(N, V), 0.0f) * hi2(inner(N, L), 0.0f) + 0.0001f;
        V3 spec = times(num, 1.0f / dn);
        radiance = plus(sh.emission, times(mulc(plus(times(mulc(kD, sh.albedo), invPi), spec), incoming), cosine));
    }
    return radiance;
}

// One iteration of the sample loop of sp_PathTraceTile, simd_path_tracer.cpp:216-320.
// Jitter: Vec2(hpw * RandomBilateral(rng), hph * RandomBilateral(rng)) -- g++ evaluates the
// second argument first, so the first draw goes to y (pinned against oracle/_ref by
// tests/test_oracle_port.py::test_port_equals_reference_image).
static V3 one_path(const ora_Scene *s, uint32_t x, uint32_t y, uint32_t *rng, uint32_t bounces,
                   float clampTo, Counters64 *m, Scratch *sc, PathVertex *path)
{
    const Camera &c = s->cam;
    float jy = c.halfPixelH * rnd11(rng);
    float jx = c.halfPixelW * rnd11(rng);
    V2 pixel = V2{((float)x + 0.5f) + jx, ((float)y + 0.5f) + jy};
    V3 filmP = film_point(c, pixel);
    V3 o = c.position;
    V3 d = unit(minus(filmP, c.position));
    uint32_t len = 0;
    for (uint32_t b = 0; b < bounces; ++b)
    {
        SceneHit h = hit_scene(s, o, d, m, sc);
        m->v[ORA_METRIC_RAYS]++;
        PathVertex *pv = path + len++;
        memset(pv, 0, sizeof(*pv));
        if (h.t > 0.0f)
        {
            pv->material = h.material;
            pv->position = plus(o, times(d, h.t));
            pv->out = flip(d);
            pv->n = h.n;
            pv->uv = h.uv;
            V3 dir = hemisphere(h.n, rng);
            pv->in = dir;
            o = plus(pv->position, times(h.n, 0.0001f));
            d = dir;
            m->v[ORA_METRIC_HITS]++;
        }
        else
        {
            pv->material = s->background;
            pv->out = flip(d);
            m->v[ORA_METRIC_MISSES]++;
            break;
        }
    }
    V3 r = path_radiance(s, path, len, clampTo);
    m->v[ORA_METRIC_PATHS]++;
    return r;
}

// sp_PathTraceTile, simd_path_tracer.cpp:178-345
static void trace_tile(const ora_Scene *s, float *rgba, uint32_t minX, uint32_t minY, uint32_t maxX,
                       uint32_t maxY, uint32_t spp, uint32_t bounces, uint32_t *rng, Counters64 *m,
                       Scratch *sc)
{
    uint64_t start = tick();
    if (maxX > s->cam.width) maxX = s->cam.width;
    if (maxY > s->cam.height) maxY = s->cam.height;
    std::vector<PathVertex> path(bounces + 1);
    for (uint32_t y = minY; y < maxY; ++y)
        for (uint32_t x = minX; x < maxX; ++x)
        {
            V3 total = {0, 0, 0};
            for (uint32_t k = 0; k < spp; ++k)
            {
                V3 r = one_path(s, x, y, rng, bounces, 10.0f, m, sc, path.data());
                total = plus(total, times(r, 1.0f / (float)spp));
            }
            float *px = rgba + ((size_t)x + (size_t)y * s->cam.width) * 4;
            px[0] = total.x; px[1] = total.y; px[2] = total.z; px[3] = 1.0f;
        }
    m->v[ORA_METRIC_CYCLES] = tick() - start;
}

// ---------------------------------------------------------------------------------------------
// harness ABI

extern "C" const char *ora_name(void) { return PORT_NAME; }
extern "C" uint32_t ora_max_bounces(void) { return 16; }
extern "C" ora_Scene *ora_create(void)
{
    ora_Scene *s = new ora_Scene();
    memset(&s->cam, 0, sizeof(s->cam));
    return s;
}
extern "C" void ora_destroy(ora_Scene *s) { delete s; }

extern "C" int ora_add_mesh(ora_Scene *s, const float *vertices, uint32_t vertexCount,
                            const uint32_t *indices, uint32_t indexCount, uint32_t smooth)
{
    Mesh mesh;
    mesh.verts.resize(vertexCount);
    memcpy(mesh.verts.data(), vertices, sizeof(Vertex) * (size_t)vertexCount);
    mesh.idx.assign(indices, indices + indexCount);
    mesh.smooth = smooth != 0;
    // sp_BuildMeshMidphase, sp_scene.cpp:21-54
    uint32_t tris = indexCount / 3;
    std::vector<V3> mn(tris), mx(tris);
    for (uint32_t i = 0; i < tris; ++i)
    {
        V3 a = mesh.verts[indices[i * 3]].p, b = mesh.verts[indices[i * 3 + 1]].p, c = mesh.verts[indices[i * 3 + 2]].p;
        mn[i] = lo3(a, lo3(b, c));
        mx[i] = hi3(a, hi3(b, c));
    }
    mesh.mid = grow_tree_any(mn.data(), mx.data(), tris);
    s->meshes.push_back(std::move(mesh));
    return (int)s->meshes.size() - 1;
}

extern "C" int ora_add_object(ora_Scene *s, uint32_t mesh, uint32_t material, const float *p,
                              const float *q, const float *sc)
{
    // sp_AddObjectToScene, sp_scene.cpp:77-117 (no 32-object cap in the port)
    const Mesh &m = s->meshes[mesh];
    V3 lo = m.verts[0].p, hi = m.verts[0].p;
    for (size_t i = 1; i < m.verts.size(); ++i) { lo = lo3(lo, m.verts[i].p); hi = hi3(hi, m.verts[i].p); }
    V3 pos = P3(p[0], p[1], p[2]), scl = P3(sc[0], sc[1], sc[2]);
    V4 rot = {q[0], q[1], q[2], q[3]};
    Object ob;
    ob.mesh = (int)mesh;
    ob.material = material;
    move_box(lo, hi, pos, rot, scl, &ob.bmin, &ob.bmax);
    ob.model = compose(compose(m_translate(pos), m_rotate(rot)), m_scale(scl));
    V3 invScale = P3(1.0f / scl.x, 1.0f / scl.y, 1.0f / scl.z);
    ob.invModel = compose(compose(m_scale(invScale), m_rotate(q_conj(rot))), m_translate(flip(pos)));
    s->objects.push_back(ob);
    return (int)s->objects.size() - 1;
}

extern "C" void ora_build(ora_Scene *s)
{
    std::vector<V3> mn(s->objects.size()), mx(s->objects.size());
    for (size_t i = 0; i < s->objects.size(); ++i) { mn[i] = s->objects[i].bmin; mx[i] = s->objects[i].bmax; }
    s->broad = grow_tree_any(mn.data(), mx.data(), (uint32_t)s->objects.size());
}

extern "C" int ora_register_material(ora_Scene *s, uint32_t id, const float *albedo,
                                     uint32_t albedoTexture, const float *emission,
                                     uint32_t emissionTexture, float roughness)
{
    if (s->mats.size() >= 32) return 0; // SP_MAX_MATERIALS, sp_material_system.h:32
    MaterialRec m = {P3(albedo[0], albedo[1], albedo[2]), albedoTexture,
                     P3(emission[0], emission[1], emission[2]), emissionTexture, roughness};
    s->matKeys.push_back(id);
    s->mats.push_back(m);
    return 1;
}

extern "C" int ora_register_texture(ora_Scene *s, uint32_t id, const float *pixels,
                                    uint32_t width, uint32_t height)
{
    if (s->imgs.size() >= 16) return 0; // SP_MAX_IMAGES
    Image img;
    img.w = width;
    img.h = height;
    img.px.assign(pixels, pixels + (size_t)width * height * 4);
    s->imgKeys.push_back(id);
    s->imgs.push_back(std::move(img));
    return 1;
}

extern "C" void ora_set_background(ora_Scene *s, uint32_t id) { s->background = id; }

extern "C" void ora_configure_camera(ora_Scene *s, const float *p, const float *q,
                                     float filmDistance, uint32_t width, uint32_t height)
{
    setup_camera(&s->cam, P3(p[0], p[1], p[2]), V4{q[0], q[1], q[2], q[3]}, filmDistance, width, height);
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

template <class Fn>
static void run_threads(uint32_t threads, Fn fn)
{
    if (threads == 0) threads = 1;
    std::vector<std::thread> pool;
    for (uint32_t t = 1; t < threads; ++t) pool.emplace_back(fn, t);
    fn(0);
    for (auto &t : pool) t.join();
}

extern "C" void ora_render_seeded(ora_Scene *s, float *rgba, uint32_t x0, uint32_t y0,
                                  uint32_t x1, uint32_t y1, uint32_t spp, uint32_t bounces,
                                  uint32_t frame, uint32_t threads, uint64_t *metrics)
{
    if (threads == 0) threads = 1;
    std::vector<Counters64> per(threads);
    memset(per.data(), 0, sizeof(Counters64) * threads);
    uint32_t width = s->cam.width;
    run_threads(threads, [&](uint32_t tid) {
        Scratch sc;
        std::vector<PathVertex> path(bounces + 1);
        Counters64 *m = &per[tid];
        float weight = 1.0f / (float)spp;
        for (uint32_t y = y0 + tid; y < y1; y += threads)
            for (uint32_t x = x0; x < x1; ++x)
            {
                V3 total = {0, 0, 0};
                for (uint32_t k = 0; k < spp; ++k)
                {
                    uint32_t rng = ora_seed(x + y * width, k, frame);
                    V3 r = one_path(s, x, y, &rng, bounces, 10.0f, m, &sc, path.data());
                    total = plus(total, times(r, weight));
                }
                float *px = rgba + ((size_t)x + (size_t)y * width) * 4;
                px[0] = total.x; px[1] = total.y; px[2] = total.z; px[3] = 1.0f;
            }
    });
    if (metrics)
        for (uint32_t t = 0; t < threads; ++t)
            for (int i = 0; i < ORA_METRIC_COUNT; ++i) metrics[i] += per[t].v[i];
}

// Pixels of the rectangle on whose paths (any sample, any bounce) a scene query ended in an exact
// tie: two triangles, or two objects, with bit-equal closest t.  There the winner depends on the
// order candidates are visited in (tree topology), which the reference does not define -- the
// parity tests accept a differing pixel only where this mask is set.  Port only.
extern "C" void ora_tie_mask(ora_Scene *s, uint8_t *mask, uint32_t x0, uint32_t y0, uint32_t x1,
                             uint32_t y1, uint32_t spp, uint32_t bounces, uint32_t frame,
                             uint32_t threads)
{
    if (threads == 0) threads = 1;
    uint32_t width = s->cam.width;
    run_threads(threads, [&](uint32_t tid) {
        Scratch sc;
        std::vector<PathVertex> path(bounces + 1);
        Counters64 m;
        memset(&m, 0, sizeof(m));
        for (uint32_t y = y0 + tid; y < y1; y += threads)
            for (uint32_t x = x0; x < x1; ++x)
            {
                uint64_t before = sc.ties;
                for (uint32_t k = 0; k < spp; ++k)
                {
                    uint32_t rng = ora_seed(x + y * width, k, frame);
                    one_path(s, x, y, &rng, bounces, 10.0f, &m, &sc, path.data());
                }
                mask[(size_t)x + (size_t)y * width] = sc.ties != before ? 1 : 0;
            }
    });
}

extern "C" double ora_render_tiles(ora_Scene *s, float *rgba, uint32_t tileW, uint32_t tileH,
                                   uint32_t spp, uint32_t bounces, uint32_t threads,
                                   uint64_t *metrics)
{
    // main.cpp:731-759,819-844: row-major tiles (tile.h:11-42) popped with an atomic counter
    // (work_queue.h:36-44), each tile seeded 0xF51C0E49
    if (threads == 0) threads = 1;
    uint32_t tilesX = (uint32_t)ceilf((float)s->cam.width / (float)tileW);
    uint32_t tilesY = (uint32_t)ceilf((float)s->cam.height / (float)tileH);
    uint32_t tileCount = tilesX * tilesY;
    std::vector<Counters64> per(tileCount);
    memset(per.data(), 0, sizeof(Counters64) * tileCount);
    int head = 0;
    auto begin = std::chrono::steady_clock::now();
    run_threads(threads, [&](uint32_t) {
        Scratch sc;
        for (;;)
        {
            int i = __atomic_fetch_add(&head, 1, __ATOMIC_SEQ_CST);
            if (i >= (int)tileCount) break;
            uint32_t tx = (uint32_t)i % tilesX, ty = (uint32_t)i / tilesX;
            uint32_t minX = tx * tileW, minY = ty * tileH;
            uint32_t maxX = minX + tileW < s->cam.width ? minX + tileW : s->cam.width;
            uint32_t maxY = minY + tileH < s->cam.height ? minY + tileH : s->cam.height;
            uint32_t rng = 0xF51C0E49u;
            trace_tile(s, rgba, minX, minY, maxX, maxY, spp, bounces, &rng, &per[i], &sc);
        }
    });
    auto end = std::chrono::steady_clock::now();
    if (metrics)
        for (uint32_t t = 0; t < tileCount; ++t)
            for (int i = 0; i < ORA_METRIC_COUNT; ++i) metrics[i] += per[t].v[i];
    return std::chrono::duration<double>(end - begin).count();
}

extern "C" double ora_render_tile_list(ora_Scene *s, float *rgba, const uint32_t *tileList,
                                       uint32_t count, uint32_t spp, uint32_t bounces,
                                       uint32_t threads, uint64_t *metrics)
{
    // as ora_render_tiles, over a caller-chosen subset of the frame's tiles
    if (threads == 0) threads = 1;
    std::vector<Counters64> per(count ? count : 1);
    memset(per.data(), 0, sizeof(Counters64) * per.size());
    int head = 0;
    auto begin = std::chrono::steady_clock::now();
    run_threads(threads, [&](uint32_t) {
        Scratch sc;
        for (;;)
        {
            int i = __atomic_fetch_add(&head, 1, __ATOMIC_SEQ_CST);
            if (i >= (int)count) break;
            const uint32_t *t = tileList + (size_t)i * 4;
            uint32_t rng = 0xF51C0E49u;
            trace_tile(s, rgba, t[0], t[1], t[2], t[3], spp, bounces, &rng, &per[i], &sc);
        }
    });
    auto end = std::chrono::steady_clock::now();
    if (metrics)
        for (uint32_t t = 0; t < count; ++t)
            for (int i = 0; i < ORA_METRIC_COUNT; ++i) metrics[i] += per[t].v[i];
    return std::chrono::duration<double>(end - begin).count();
}

extern "C" void ora_path_trace_tile(ora_Scene *s, float *rgba, uint32_t minX, uint32_t minY,
                                    uint32_t maxX, uint32_t maxY, uint32_t spp,
                                    uint32_t bounces, uint32_t *rngState, uint64_t *metrics)
{
    Scratch sc;
    Counters64 m;
    memset(&m, 0, sizeof(m));
    trace_tile(s, rgba, minX, minY, maxX, maxY, spp, bounces, rngState, &m, &sc);
    if (metrics) for (int i = 0; i < ORA_METRIC_COUNT; ++i) metrics[i] += m.v[i];
}

extern "C" void ora_primary_hits(ora_Scene *s, int32_t *triId, int32_t *objId, float *tOut,
                                 float *rayDir3, uint32_t sample, uint32_t frame,
                                 uint32_t threads)
{
    uint32_t width = s->cam.width, height = s->cam.height;
    if (threads == 0) threads = 1;
    run_threads(threads, [&](uint32_t tid) {
        Scratch sc;
        Counters64 m;
        memset(&m, 0, sizeof(m));
        for (uint32_t y = tid; y < height; y += threads)
            for (uint32_t x = 0; x < width; ++x)
            {
                uint32_t rng = ora_seed(x + y * width, sample, frame);
                const Camera &c = s->cam;
                float jy = c.halfPixelH * rnd11(&rng);
                float jx = c.halfPixelW * rnd11(&rng);
                V3 filmP = film_point(c, V2{((float)x + 0.5f) + jx, ((float)y + 0.5f) + jy});
                V3 d = unit(minus(filmP, c.position));
                SceneHit h = hit_scene(s, c.position, d, &m, &sc);
                uint32_t i = x + y * width;
                if (triId) triId[i] = h.triangle;
                if (objId) objId[i] = h.object;
                if (tOut) tOut[i] = h.t;
                if (rayDir3) { rayDir3[i * 3] = d.x; rayDir3[i * 3 + 1] = d.y; rayDir3[i * 3 + 2] = d.z; }
            }
    });
}

extern "C" void ora_intersect_rays(ora_Scene *s, uint32_t n, const float *origins3,
                                   const float *dirs3, float *out7, int32_t *triId,
                                   int32_t *objId, uint64_t *metrics)
{
    Scratch sc;
    Counters64 m;
    memset(&m, 0, sizeof(m));
    for (uint32_t i = 0; i < n; ++i)
    {
        V3 o = P3(origins3[i * 3], origins3[i * 3 + 1], origins3[i * 3 + 2]);
        V3 d = P3(dirs3[i * 3], dirs3[i * 3 + 1], dirs3[i * 3 + 2]);
        SceneHit h = hit_scene(s, o, d, &m, &sc);
        if (out7)
        {
            float *out = out7 + (size_t)i * 7;
            out[0] = h.t;
            memcpy(&out[1], &h.material, 4);
            out[2] = h.n.x; out[3] = h.n.y; out[4] = h.n.z;
            out[5] = h.uv.x; out[6] = h.uv.y;
        }
        if (triId) triId[i] = h.triangle;
        if (objId) objId[i] = h.object;
    }
    if (metrics) for (int i = 0; i < ORA_METRIC_COUNT; ++i) metrics[i] += m.v[i];
}

// ---- known-answer entry points ----

extern "C" uint32_t ora_xorshift32(uint32_t *state) { return xs32(state); }
extern "C" float ora_random_unilateral(uint32_t *state) { return rnd01(state); }
extern "C" float ora_random_bilateral(uint32_t *state) { return rnd11(state); }

extern "C" void ora_ray_triangle_mt(const float *o, const float *d, const float *a,
                                    const float *b, const float *c, float *out6)
{
    TriHit r = moller_trumbore(P3(o[0], o[1], o[2]), P3(d[0], d[1], d[2]), P3(a[0], a[1], a[2]),
                               P3(b[0], b[1], b[2]), P3(c[0], c[1], c[2]));
    out6[0] = r.t; out6[1] = r.uv.x; out6[2] = r.uv.y; out6[3] = r.n.x; out6[4] = r.n.y; out6[5] = r.n.z;
}

extern "C" uint32_t ora_ray_aabb4(const float *boxMin12, const float *boxMax12, const float *o,
                                  const float *inv)
{
    uint32_t mask = 0;
    for (int i = 0; i < 4; ++i)
        if (slab_lane(P3(boxMin12[i * 3], boxMin12[i * 3 + 1], boxMin12[i * 3 + 2]),
                      P3(boxMax12[i * 3], boxMax12[i * 3 + 1], boxMax12[i * 3 + 2]),
                      P3(o[0], o[1], o[2]), P3(inv[0], inv[1], inv[2])))
            mask |= 1u << i;
    return mask;
}

extern "C" float ora_ray_aabb_scalar(const float *mn, const float *mx, const float *o,
                                     const float *d)
{
    return slab_scalar(P3(mn[0], mn[1], mn[2]), P3(mx[0], mx[1], mx[2]), P3(o[0], o[1], o[2]), P3(d[0], d[1], d[2]));
}

extern "C" void ora_hemisphere(uint32_t *state, const float *n, float *out3)
{
    V3 v = hemisphere(P3(n[0], n[1], n[2]), state);
    out3[0] = v.x; out3[1] = v.y; out3[2] = v.z;
}

extern "C" void ora_to_spherical(const float *v, float *out2) { V2 r = to_sphere(P3(v[0], v[1], v[2])); out2[0] = r.x; out2[1] = r.y; }
extern "C" void ora_map_equirect(const float *sc, float *out2) { V2 r = to_equirect(V2{sc[0], sc[1]}); out2[0] = r.x; out2[1] = r.y; }
extern "C" void ora_spherical_to_cartesian(const float *sc, float *out3) { V3 r = from_sphere(V2{sc[0], sc[1]}); out3[0] = r.x; out3[1] = r.y; out3[2] = r.z; }

extern "C" void ora_camera_fields(const float *p, const float *q, float filmDistance,
                                  uint32_t width, uint32_t height, float *out)
{
    Camera c;
    setup_camera(&c, P3(p[0], p[1], p[2]), V4{q[0], q[1], q[2], q[3]}, filmDistance, width, height);
    V3 v[5] = {c.right, c.up, c.forward, c.position, c.filmCenter};
    for (int i = 0; i < 5; ++i) { out[i * 3] = v[i].x; out[i * 3 + 1] = v[i].y; out[i * 3 + 2] = v[i].z; }
    out[15] = c.halfPixelW; out[16] = c.halfPixelH; out[17] = c.halfFilmW; out[18] = c.halfFilmH;
    out[19] = out[20] = out[21] = 0.0f;
}

extern "C" void ora_film_positions(ora_Scene *s, uint32_t n, const float *pixelPos2, float *out3)
{
    for (uint32_t i = 0; i < n; ++i)
    {
        V3 f = film_point(s->cam, V2{pixelPos2[i * 2], pixelPos2[i * 2 + 1]});
        out3[i * 3] = f.x; out3[i * 3 + 1] = f.y; out3[i * 3 + 2] = f.z;
    }
}

extern "C" void ora_transform_aabb(const float *mn, const float *mx, const float *p,
                                   const float *q, const float *sc, float *out6)
{
    V3 lo, hi;
    move_box(P3(mn[0], mn[1], mn[2]), P3(mx[0], mx[1], mx[2]), P3(p[0], p[1], p[2]),
             V4{q[0], q[1], q[2], q[3]}, P3(sc[0], sc[1], sc[2]), &lo, &hi);
    out6[0] = lo.x; out6[1] = lo.y; out6[2] = lo.z; out6[3] = hi.x; out6[4] = hi.y; out6[5] = hi.z;
}

extern "C" void ora_radiance_for_path(ora_Scene *s, const float *path15, uint32_t n, float *out3)
{
    std::vector<PathVertex> path(n ? n : 1);
    for (uint32_t i = 0; i < n; ++i)
    {
        const float *p = path15 + (size_t)i * 15;
        PathVertex v;
        memcpy(&v.material, &p[0], 4);
        v.position = P3(p[1], p[2], p[3]);
        v.out = P3(p[4], p[5], p[6]);
        v.in = P3(p[7], p[8], p[9]);
        v.n = P3(p[10], p[11], p[12]);
        v.uv = V2{p[13], p[14]};
        path[i] = v;
    }
    V3 r = path_radiance(s, path.data(), n, 10.0f);
    out3[0] = r.x; out3[1] = r.y; out3[2] = r.z;
}

// post_processing.frag.glsl:19-26 (PerformToneMapping) + UNORM8 store; pow follows the libm switch
// of this file (powf, or pow in double rounded once under ORA_DETERMINISTIC_MATH)
extern "C" void ora_tone_map(const float *rgba, uint32_t count, float exposure, uint32_t *out)
{
    for (uint32_t i = 0; i < count; ++i)
    {
        uint32_t packed = 0xFF000000u;
        for (int ch = 0; ch < 3; ++ch)
        {
            float c = rgba[(size_t)i * 4 + ch];
            c = c * exposure;
            c = c / (1.0f + c);
#ifdef ORA_DETERMINISTIC_MATH
            c = (float)pow((double)c, 1.0 / 2.2);
#else
            c = powf(c, 1.0f / 2.2f);
#endif
            float q = c > 0.0f ? (c < 1.0f ? c : 1.0f) : 0.0f; // NaN -> 0
            uint32_t byte = (uint32_t)floorf(q * 255.0f + 0.5f);
            packed |= byte << (8 * ch);
        }
        out[i] = packed;
    }
}

extern "C" void ora_sample_nearest(const float *pixels, uint32_t w, uint32_t h, float u, float v, float *out4)
{
    Image img;
    img.w = w; img.h = h;
    img.px.assign(pixels, pixels + (size_t)w * h * 4);
    V4 r = nearest(img, V2{u, v});
    out4[0] = r.x; out4[1] = r.y; out4[2] = r.z; out4[3] = r.w;
}

extern "C" void ora_sample_bilinear(const float *pixels, uint32_t w, uint32_t h, float u, float v, float *out4)
{
    Image img;
    img.w = w; img.h = h;
    img.px.assign(pixels, pixels + (size_t)w * h * 4);
    V4 r = bilinear(img, V2{u, v});
    out4[0] = r.x; out4[1] = r.y; out4[2] = r.z; out4[3] = r.w;
}

// ---------------------------------------------------------------------------------------------
// Environment pre-processing (src/cubemap.cpp), restated.  Faces +X -X +Y -Y +Z -Z.

// MapCubeMapLayerIndexToBasisVectors, cubemap.cpp:54-104: {forward, up, right} per layer
static void face_axes(uint32_t layer, V3 *fwd, V3 *up, V3 *right)
{
    static const float table[6][9] = {
        {1, 0, 0, 0, 1, 0, 0, 0, -1},  {-1, 0, 0, 0, 1, 0, 0, 0, 1}, {0, 1, 0, 0, 0, -1, 1, 0, 0},
        {0, -1, 0, 0, 0, 1, 1, 0, 0},  {0, 0, 1, 0, 1, 0, 1, 0, 0},  {0, 0, -1, 0, 1, 0, -1, 0, 0}};
    const float *t = table[layer];
    *fwd = P3(t[0], t[1], t[2]);
    *up = P3(t[3], t[4], t[5]);
    *right = P3(t[6], t[7], t[8]);
}

// cubemap.cpp:263-275 (and :139-150)
static V3 face_direction(V3 fwd, V3 up, V3 right, uint32_t x, uint32_t y, uint32_t w, uint32_t h)
{
    float fx = (float)x / (float)w, fy = (float)y / (float)h;
    fy = 1.0f - fy;
    fx = fx * 2.0f - 1.0f;
    fy = fy * 2.0f - 1.0f;
    return unit(plus(plus(fwd, times(right, fx)), times(up, fy)));
}

// cubemap.cpp:277-281: direction -> spherical -> equirect uv, v flipped -> bilinear
static V4 env_bilinear(const Image &env, V3 d)
{
    V2 uv = to_equirect(to_sphere(d));
    uv.y = 1.0f - uv.y;
    return bilinear(env, uv);
}

static inline float clamp_to(float v, float c) { return lo2(hi2(v, 0.0f), c); } // Clamp, math_utils.h:135-139

static Image wrap_image(const float *pixels, uint32_t w, uint32_t h)
{
    Image img;
    img.w = w;
    img.h = h;
    img.px.assign(pixels, pixels + (size_t)w * h * 4);
    return img;
}

// CreateCubeMap, cubemap.cpp:237-291
extern "C" void ora_create_cube_map(const float *pixels, uint32_t w, uint32_t h, uint32_t faceW,
                                    uint32_t faceH, float *out)
{
    Image env = wrap_image(pixels, w, h);
    for (uint32_t layer = 0; layer < 6; ++layer)
    {
        V3 fwd, up, right;
        face_axes(layer, &fwd, &up, &right);
        for (uint32_t y = 0; y < faceH; ++y)
            for (uint32_t x = 0; x < faceW; ++x)
            {
                V4 s = env_bilinear(env, face_direction(fwd, up, right, x, y, faceW, faceH));
                float *dst = out + (((size_t)layer * faceH + y) * faceW + x) * 4;
                dst[0] = s.x; dst[1] = s.y; dst[2] = s.z; dst[3] = s.w;
            }
    }
}

// CreateIrradianceCubeMap, cubemap.cpp:108-233.  sampling 0: the uniform (phi, theta) grid
// (:152-198, the branch config.h:47 compiles); sampling 1: random offsets from one serial stream
// (:200-224).  sampleDelta is the 0.1f of :160.
extern "C" int ora_create_irradiance_cube_map(const float *pixels, uint32_t w, uint32_t h, uint32_t faceW,
                                               uint32_t faceH, uint32_t samplesPerPixel, uint32_t sampling,
                                               float sampleDelta, float *out)
{
    Image env = wrap_image(pixels, w, h);
    const float clampValue = 10.0f; // RADIANCE_CLAMP, config.h:36
    float contribution = 1.0f / (float)samplesPerPixel;
    uint32_t rng = 0x45BA12F3u; // :123, never reseeded
    for (uint32_t layer = 0; layer < 6; ++layer)
    {
        V3 fwd, up, right;
        face_axes(layer, &fwd, &up, &right);
        for (uint32_t y = 0; y < faceH; ++y)
            for (uint32_t x = 0; x < faceW; ++x)
            {
                V3 dir = face_direction(fwd, up, right, x, y, faceW, faceH);
                V3 sum = P3(0, 0, 0);
                if (sampling == 0)
                {
                    V3 tangent = unit(outer(up, dir));
                    V3 bitangent = unit(outer(dir, tangent));
                    uint32_t count = 0;
                    for (float phi = 0.0f; phi < 2.0f * kPi; phi += sampleDelta)
                        for (float theta = 0.0f; theta < 0.5f * kPi; theta += sampleDelta)
                        {
                            V3 local = from_sphere(V2{phi, theta});
                            V3 world = plus(plus(times(dir, local.y), times(tangent, local.x)),
                                            times(bitangent, local.z));
                            V4 s = env_bilinear(env, world);
                            V3 radiance = P3(clamp_to(s.x, clampValue), clamp_to(s.y, clampValue),
                                             clamp_to(s.z, clampValue));
                            sum = plus(sum, times(times(radiance, lm_cos(theta)), lm_sin(theta)));
                            count++;
                        }
                    sum = times(times(sum, kPi), 1.0f / (float)count);
                }
                else
                {
                    for (uint32_t i = 0; i < samplesPerPixel; ++i)
                    {
                        // Vec3(RandomBilateral, RandomBilateral, RandomBilateral): g++ evaluates
                        // call arguments right to left, so z is drawn first
                        float oz = rnd11(&rng);
                        float oy = rnd11(&rng);
                        float ox = rnd11(&rng);
                        V3 sd = unit(plus(dir, P3(ox, oy, oz)));
                        if (inner(sd, dir) < 0.0f) sd = flip(sd);
                        float cosine = hi2(inner(dir, sd), 0.0f);
                        V4 s = env_bilinear(env, sd);
                        V3 radiance = P3(clamp_to(s.x * cosine, clampValue), clamp_to(s.y * cosine, clampValue),
                                         clamp_to(s.z * cosine, clampValue));
                        sum = plus(sum, times(radiance, contribution));
                    }
                }
                float *dst = out + (((size_t)layer * faceH + y) * faceW + x) * 4;
                dst[0] = sum.x; dst[1] = sum.y; dst[2] = sum.z; dst[3] = 1.0f;
            }
    }
    return 1;
}

// ComputeTiles, tile.h:11-42
extern "C" uint32_t ora_compute_tiles(uint32_t w, uint32_t h, uint32_t tw, uint32_t th,
                                      uint32_t *tiles, uint32_t maxTiles)
{
    uint32_t ny = (uint32_t)ceilf((float)h / (float)th), nx = (uint32_t)ceilf((float)w / (float)tw);
    for (uint32_t ty = 0; ty < ny; ++ty)
        for (uint32_t tx = 0; tx < nx; ++tx)
        {
            uint32_t i = tx + ty * nx;
            if (i >= maxTiles) break;
            uint32_t x = tx * tw, y = ty * th;
            tiles[i * 4 + 0] = x;
            tiles[i * 4 + 1] = y;
            tiles[i * 4 + 2] = x + tw < w ? x + tw : w;
            tiles[i * 4 + 3] = y + th < h ? y + th : h;
        }
    uint32_t total = ny * nx;
    return total < maxTiles ? total : maxTiles;
}

extern "C" uint32_t ora_bvh_query(const float *aabbMin, const float *aabbMax, uint32_t n,
                                  const float *o, const float *d, uint32_t *leaves,
                                  uint32_t maxLeaves, uint32_t *error, uint32_t *aabbTests,
                                  float *rootBounds6)
{
    Tree tree = grow_tree((const V3 *)aabbMin, (const V3 *)aabbMax, n);
    std::vector<int> out, st[2];
    // cap semantics of bvh.cpp:288-302: with cap == 0 slots nothing can be stored
    Query q;
    if (maxLeaves == 0)
    {
        q = walk_tree(tree, P3(o[0], o[1], o[2]), P3(d[0], d[1], d[2]), out, 0, st);
        q.overflow = q.count > 0;
        q.count = 0;
    }
    else
        q = walk_tree(tree, P3(o[0], o[1], o[2]), P3(d[0], d[1], d[2]), out, maxLeaves, st);
    for (uint32_t i = 0; i < q.count; ++i) leaves[i] = tree.nodes[out[i]].leaf;
    if (error) *error = q.overflow ? 1 : 0;
    if (aabbTests) *aabbTests = q.aabbTests;
    if (rootBounds6 && tree.root >= 0)
    {
        const TreeNode &r = tree.nodes[tree.root];
        rootBounds6[0] = r.mn.x; rootBounds6[1] = r.mn.y; rootBounds6[2] = r.mn.z;
        rootBounds6[3] = r.mx.x; rootBounds6[4] = r.mx.y; rootBounds6[5] = r.mx.z;
    }
    return q.count;
}

// perf_tests/perf_tests.cpp:51-118 (TestBvh) and :212-305 (TestMeshMidphase) as calls, single-threaded
extern "C" double ora_perf_bvh(const float *aabbMin, const float *aabbMax, uint32_t n, uint32_t rays, const float *origins3,
                               const float *dirs3, uint32_t maxLeaves, uint32_t *countXorSum, double *buildSeconds)
{
    auto b0 = std::chrono::steady_clock::now();
    Tree tree = grow_tree((const V3 *)aabbMin, (const V3 *)aabbMax, n);
    if (buildSeconds) *buildSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - b0).count();
    std::vector<int> out, st[2];
    auto t0 = std::chrono::steady_clock::now();
    for (uint32_t q = 0; q < rays; ++q)
    {
        Query r = walk_tree(tree, P3(origins3[q * 3], origins3[q * 3 + 1], origins3[q * 3 + 2]),
                            P3(dirs3[q * 3], dirs3[q * 3 + 1], dirs3[q * 3 + 2]), out, maxLeaves, st);
        uint32_t x = 0, sum = 0;
        for (uint32_t i = 0; i < r.count; ++i)
        {
            uint32_t leaf = (uint32_t)tree.nodes[out[i]].leaf;
            x ^= leaf * 0x9E3779B1u;
            sum += leaf;
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
    Counters64 m;
    memset(&m, 0, sizeof(m));
    Scratch sc;
    const Mesh &me = s->meshes[mesh];
    auto t0 = std::chrono::steady_clock::now();
    for (uint32_t q = 0; q < rays; ++q)
    {
        MeshHit h = hit_mesh(me, P3(origins3[q * 3], origins3[q * 3 + 1], origins3[q * 3 + 2]),
                             P3(dirs3[q * 3], dirs3[q * 3 + 1], dirs3[q * 3 + 2]), &m, &sc);
        t[q] = h.tri.t;
        if (tri) tri[q] = h.triangle;
    }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

static void tree_stats(const Tree &t, int ni, uint32_t depth, uint32_t *leaves, uint32_t *internal,
                       uint32_t *minDepth, uint32_t *maxDepth, uint32_t *contained, std::vector<uint8_t> *seen)
{
    const TreeNode &n = t.nodes[ni];
    if (n.child[0] < 0)
    {
        (*leaves)++;
        if (depth < *minDepth) *minDepth = depth;
        if (depth > *maxDepth) *maxDepth = depth;
        if (n.leaf < seen->size()) (*seen)[n.leaf] = 1;
        return;
    }
    (*internal)++;
    for (int i = 0; i < 4; ++i)
    {
        if (n.child[i] < 0) continue;
        const TreeNode &c = t.nodes[n.child[i]];
        if (c.mn.x < n.mn.x || c.mn.y < n.mn.y || c.mn.z < n.mn.z || c.mx.x > n.mx.x || c.mx.y > n.mx.y || c.mx.z > n.mx.z)
            *contained = 0;
        tree_stats(t, n.child[i], depth + 1, leaves, internal, minDepth, maxDepth, contained, seen);
    }
}

extern "C" void ora_mesh_tree_stats(ora_Scene *s, uint32_t mesh, uint32_t *out6)
{
    const Mesh &m = s->meshes[mesh];
    uint32_t leaves = 0, internal = 0, minDepth = 0xFFFFFFFFu, maxDepth = 0, contained = 1;
    std::vector<uint8_t> seen(m.idx.size() / 3, 0);
    if (m.mid.root >= 0) tree_stats(m.mid, m.mid.root, 0, &leaves, &internal, &minDepth, &maxDepth, &contained, &seen);
    uint32_t all = 1;
    for (uint8_t b : seen) if (!b) all = 0;
    out6[0] = leaves; out6[1] = internal; out6[2] = minDepth; out6[3] = maxDepth; out6[4] = all; out6[5] = contained;
}
