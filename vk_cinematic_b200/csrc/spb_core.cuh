// spb_core.cuh -- the arithmetic of the sp_ path, written once for the device.
//
// Every function here states the reference lines whose *operation order* it reproduces; the
// translation units that include this header are compiled with -fmad=false (nvcc) so no
// multiply-add is contracted, and IEEE division / square root (nvcc defaults).  The header also
// compiles as plain C++ (g++ -ffp-contract=off) for tests/hostsim, a debugging aid that lets the
// traversal and shading logic be checked bit-for-bit against the oracle without a GPU.  The
// shipped library never runs this code on the host.
#pragma once

#include <stdint.h>
#include <string.h>
#include <math.h>
#include <float.h>

#if defined(__CUDACC__)
// (out of line on purpose: optional paths whose registers and local arrays must not be charged to
// the kernels' hot loops)
#define SPB_HD_NOINLINE static __host__ __device__ __noinline__
#define SPB_HD __host__ __device__ __forceinline__
#define SPB_ALIGN16 __align__(16)
#define SPB_ALIGN8 __align__(8)
#else
#define SPB_HD_NOINLINE inline
#define SPB_HD inline
#define SPB_ALIGN16 alignas(16)
#define SPB_ALIGN8 alignas(8)
#endif

namespace spb {

#if defined(__CUDACC__)
typedef float4 v4f;
typedef uint4 v4u;
#else
struct SPB_ALIGN16 v4f { float x, y, z, w; };
struct SPB_ALIGN16 v4u { uint32_t x, y, z, w; };
#endif

SPB_HD v4f ld4(const v4f *p)
{
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}
SPB_HD v4u ld4u(const v4u *p)
{
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}
// Pull a line towards L1 ahead of use (a popped node is fetched by 32 lanes at once and the
// slowest lane sets the pace, so the expected miss is issued as soon as the address is known).
SPB_HD void prefetch_l1(const void *p)
{
#if defined(__CUDA_ARCH__) && defined(SPB_PREFETCH)
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}
// 32-byte load through the read-only path (LDG.E.256 on sm_100a): half as many L1 wavefronts per
// node as 16-byte loads when every lane of a warp reads a different node.
SPB_HD void ld8(const v4f *p, v4f &a, v4f &b)
{
#if defined(__CUDA_ARCH__) && !defined(SPB_NO_LD256)
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                 : "l"(p));
#else
    a = ld4(p);
    b = ld4(p + 1);
#endif
}
SPB_HD uint32_t f2u(float f)
{
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
#endif
}
SPB_HD float u2f(uint32_t u)
{
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}

// ------------------------------------------------------------------------------------------
// constants (math_utils.h:6-7)
#define SPB_PI 3.14159265359f
#define SPB_EPSILON FLT_EPSILON

#define SPB_REF_EMPTY 0xFFFFFFFFu
#define SPB_REF_LEAF 0x80000000u
#define SPB_STACK_SIZE 96
// Shares of the traversal stack: intersect_scene() gives the TLAS walk the first third and the
// mesh walk the rest; the resumable machine uses one stack for both plus two sentinel entries.
// build_bvh4() keeps every tree inside its share (median-split rebuild), flatten_scene() and
// upload_scene() check it.
#define SPB_TLAS_STACK_LIMIT (SPB_STACK_SIZE / 3 - 2)
#define SPB_MESH_STACK_LIMIT ((SPB_STACK_SIZE * 2) / 3 - 4)
#define SPB_MAX_BOUNCES 8
#define SPB_MAX_MATERIALS 32
#define SPB_MAX_IMAGES 16

// ------------------------------------------------------------------------------------------
// vec3 with the reference's evaluation order (math_lib.h:178-215, 266-286, 461-521)
struct f3 { float x, y, z; };

SPB_HD f3 mk3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
SPB_HD f3 add3(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
SPB_HD f3 sub3(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
SPB_HD f3 mul3(f3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
SPB_HD f3 neg3(f3 a) { return mk3(-a.x, -a.y, -a.z); }
SPB_HD f3 had3(f3 a, f3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
SPB_HD float dot3(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
SPB_HD f3 cross3(f3 a, f3 b)
{
    return mk3((a.y * b.z) - (a.z * b.y), (a.z * b.x) - (a.x * b.z), (a.x * b.y) - (a.y * b.x));
}
// Min/Max as `a < b ? a : b` / `a > b ? a : b` (math_utils.h:9-19): NaN picks b.
SPB_HD float rmin(float a, float b) { return a < b ? a : b; }
SPB_HD float rmax(float a, float b) { return a > b ? a : b; }

// Normalize (math_lib.h:503-512): zero vector unless length > FLT_EPSILON; multiply by reciprocal
SPB_HD f3 normalize3(f3 v, float *lengthOut = nullptr)
{
    float length = sqrtf(dot3(v, v));
    if (lengthOut) *lengthOut = length;
    f3 r = mk3(0.0f, 0.0f, 0.0f);
    if (length > SPB_EPSILON)
    {
        r = mul3(v, 1.0f / length);
    }
    return r;
}

// mat4 * vec4 (math_lib.h:381-398 with Dot at :238-248): columns m[0..3]; row i dotted with
// (v, w) left to right, including the w term (it can flip the sign of a zero).
struct m4 { v4f c[4]; };
SPB_HD f3 xform(const m4 &m, f3 v, float w)
{
    f3 r;
    r.x = m.c[0].x * v.x + m.c[1].x * v.y + m.c[2].x * v.z + m.c[3].x * w;
    r.y = m.c[0].y * v.x + m.c[1].y * v.y + m.c[2].y * v.z + m.c[3].y * w;
    r.z = m.c[0].z * v.x + m.c[1].z * v.y + m.c[2].z * v.z + m.c[3].z * w;
    return r;
}

// ------------------------------------------------------------------------------------------
// libm stand-ins.  MATH == 0: evaluate in double and round once to float -- agrees with glibc's
// sinf/cosf/atan2f/powf wherever those are correctly rounded (98.7 % / 92 % / 99.93 % of inputs,
// measured; DESIGN.md "libm") and is what the oracle's deterministic-math mode computes.
// MATH == 1: CUDA's single-precision functions.
template <int MATH> SPB_HD float m_sin(float x) { return MATH == 0 ? (float)sin((double)x) : sinf(x); }
template <int MATH> SPB_HD float m_cos(float x) { return MATH == 0 ? (float)cos((double)x) : cosf(x); }
template <int MATH> SPB_HD float m_atan2(float y, float x)
{
    return MATH == 0 ? (float)atan2((double)y, (double)x) : atan2f(y, x);
}
// Pow(x, 5.0f) (simd_path_tracer.cpp:65-70): x^5 by three double multiplies, one rounding
template <int MATH> SPB_HD float m_pow5(float x)
{
    if (MATH == 0)
    {
        double d = (double)x;
        double d2 = d * d;
        return (float)(d2 * d2 * d);
    }
    return powf(x, 5.0f);
}

// ------------------------------------------------------------------------------------------
// RNG (math_utils.h:184-214)
SPB_HD uint32_t xorshift32(uint32_t &state)
{
    uint32_t x = state;
    x ^= x << 13;
    x ^= x >> 17;
    x ^= x << 5;
    state = x;
    return x;
}
SPB_HD float rand_unilateral(uint32_t &state)
{
    // (f32)(x >> 1) / (f32)(U32_MAX >> 1); the denominator rounds to 2^31
    return (float)(xorshift32(state) >> 1) / 2147483648.0f;
}
SPB_HD float rand_bilateral(uint32_t &state) { return -1.0f + 2.0f * rand_unilateral(state); }

// Seed of the (pixel, sample, frame) stream; never 0 (XorShift32's fixed point).
SPB_HD uint32_t stream_seed(uint32_t pixelIndex, uint32_t sample, uint32_t frame)
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

// ------------------------------------------------------------------------------------------
// device-side scene description (built and owned by spb_host.cpp)
struct DCamera
{
    f3 right, up, position, filmCenter;
    float halfPixelWidth, halfPixelHeight, halfFilmWidth, halfFilmHeight;
    uint32_t width, height;
};

struct DImage { const v4f *pixels; uint32_t width, height, pad; };

struct DMaterials
{
    uint32_t count;
    uint32_t backgroundId;
    uint32_t imageCount;
    // 1 when the vertex terms of an escaped ray cannot influence its radiance (see
    // miss_radiance()): set by convert_materials()
    uint32_t simpleBackground;
    uint32_t keys[SPB_MAX_MATERIALS];
    float albedo[SPB_MAX_MATERIALS][3];
    float emission[SPB_MAX_MATERIALS][3];
    float roughness[SPB_MAX_MATERIALS];
    int32_t albedoImage[SPB_MAX_MATERIALS];   // resolved sp_FindTexture result, -1 = none
    int32_t emissionImage[SPB_MAX_MATERIALS];
    DImage images[SPB_MAX_IMAGES];
};

// Node = 8 x 16 B: child boxes SoA (min x,y,z then max x,y,z, four lanes each), 4 child refs,
// 16 B of metadata.  A child is either another node, one primitive (SPB_REF_LEAF | slot) whose
// box is the primitive's own AABB -- which makes the parent's slab test the reference's
// per-leaf test (bvh.cpp:236-255) -- or SPB_REF_EMPTY.
struct DScene
{
    const v4f *nodes;      // 8 per node
    const v4f *tris;       // 3 per triangle slot: (v0, bits triIndex) (v1, -) (v2, -)
    const v4f *shade;      // 4 per triangle: (n0,uv0.x) (n1,uv0.y) (n2,uv1.x) (uv1.y,uv2.x,uv2.y,-)
    const v4f *objInv;     // 4 per object: inverse model matrix columns
    const v4f *objModel;   // 4 per object: model matrix columns
    const v4u *objInfo;    // x = mesh root node (or EMPTY), y = shade base, z = smooth flag, w = material
    const uint32_t *objTris; // per object: first triangle slot of its mesh; then objectCount + 1 prefix
                             // sums of the objects' triangle counts (coverage pass)
    const v4f *objBox;     // 2 per object: (world AABB min, largest |coordinate| of its mesh tree's boxes)
                           // (world AABB max, -): the box the TLAS leaf carries (sp_scene.cpp:98-116), kept
                           // apart so that a traversal that reaches the object through a CONSERVATIVE
                           // inner test can still apply the reference's exact test to the object's own box
    uint32_t tlasRoot;     // node index or SPB_REF_EMPTY
    uint32_t objectCount;
    float tlasExtent;      // largest |coordinate| of the TLAS boxes (world space)
    uint32_t tlasNodeCount; // TLAS nodes are [tlasRoot, tlasRoot + tlasNodeCount), breadth-first
    uint32_t triangleTest; // 0: the reference's Moller-Trumbore (parity); 1: watertight (sp_b200_Params::triangleTest)
};

struct Counters
{
    uint32_t nodeVisits, triangleTests, objectTests, envClamped;
};

// ------------------------------------------------------------------------------------------
// camera (simd_path_tracer.cpp:38-63, 216-230)
SPB_HD void primary_ray(const DCamera &cam, uint32_t x, uint32_t y, uint32_t &rng, f3 &origin,
                        f3 &direction)
{
    // Vec2(halfPixelWidth * RandomBilateral(rng), halfPixelHeight * RandomBilateral(rng)):
    // g++ evaluates the arguments right to left, so the first draw jitters y (pinned by
    // tests/test_oracle_ref.py::test_jitter_draw_order).
    float by = rand_bilateral(rng);
    float bx = rand_bilateral(rng);
    float px = ((float)x + 0.5f) + cam.halfPixelWidth * bx;
    float py = ((float)y + 0.5f) + cam.halfPixelHeight * by;

    float fx = px / (float)cam.width;
    float fy = py / (float)cam.height;
    fy = 1.0f - fy;
    fx = fx * 2.0f - 1.0f;
    fy = fy * 2.0f - 1.0f;
    f3 filmP = mul3(cam.right, cam.halfFilmWidth * fx);
    filmP = add3(filmP, mul3(cam.up, cam.halfFilmHeight * fy));
    filmP = add3(filmP, cam.filmCenter);

    origin = cam.position;
    direction = normalize3(sub3(filmP, cam.position));
}

// ------------------------------------------------------------------------------------------
// slab test of one lane of a 4-wide node (simd.h:198-271): per axis t0 = (min-o)*inv,
// t1 = (max-o)*inv, tmin = minps(t0,t1), tmax = maxps(t0,t1) (second operand on NaN), then
// Max(0, max3(tmin)) <= min3(tmax) with the scalar Max/Min.  tnear is the left-hand side.
SPB_HD bool slab_exact(float bminx, float bminy, float bminz, float bmaxx, float bmaxy,
                       float bmaxz, f3 o, f3 inv, float &tnear)
{
    float t0x = (bminx - o.x) * inv.x, t1x = (bmaxx - o.x) * inv.x;
    float t0y = (bminy - o.y) * inv.y, t1y = (bmaxy - o.y) * inv.y;
    float t0z = (bminz - o.z) * inv.z, t1z = (bmaxz - o.z) * inv.z;
    float tminx = rmin(t0x, t1x), tmaxx = rmax(t0x, t1x);
    float tminy = rmin(t0y, t1y), tmaxy = rmax(t0y, t1y);
    float tminz = rmin(t0z, t1z), tmaxz = rmax(t0z, t1z);
    tnear = rmax(0.0f, rmax(tminx, rmax(tminy, tminz)));
    float tfar = rmin(tmaxx, rmin(tmaxy, tmaxz));
    return tnear <= tfar;
}
// Same predicate when no product can be NaN (all inv components finite): hardware min/max.
SPB_HD bool slab_fast(float bminx, float bminy, float bminz, float bmaxx, float bmaxy,
                      float bmaxz, f3 o, f3 inv, float &tnear)
{
    float t0x = (bminx - o.x) * inv.x, t1x = (bmaxx - o.x) * inv.x;
    float t0y = (bminy - o.y) * inv.y, t1y = (bmaxy - o.y) * inv.y;
    float t0z = (bminz - o.z) * inv.z, t1z = (bmaxz - o.z) * inv.z;
    tnear = fmaxf(0.0f, fmaxf(fminf(t0x, t1x), fmaxf(fminf(t0y, t1y), fminf(t0z, t1z))));
    float tfar = fminf(fmaxf(t0x, t1x), fminf(fmaxf(t0y, t1y), fmaxf(t0z, t1z)));
    return tnear <= tfar;
}

// Watertight mode (sp_b200_Params::triangleTest): the box tests must not lose a ray the watertight
// triangle test would accept -- a ray that grazes the common edge of two triangles also grazes a face
// of both their boxes, where the slab test's rounding can reject it twice.  Every box is therefore
// grown, per axis, by 2^-20 of the magnitudes involved before it is tested (Woop et al. 2013 section 4
// pad their boxes for the same reason).  NaN boxes (empty slots) stay NaN.  Not used in parity mode.
SPB_HD void grow_box(f3 o, float &bminx, float &bminy, float &bminz, float &bmaxx, float &bmaxy, float &bmaxz)
{
    const float e = 9.5367431640625e-07f; // 2^-20
    const float px = e * (fabsf(o.x) + fmaxf(fabsf(bminx), fabsf(bmaxx)));
    const float py = e * (fabsf(o.y) + fmaxf(fabsf(bminy), fabsf(bmaxy)));
    const float pz = e * (fabsf(o.z) + fmaxf(fabsf(bminz), fabsf(bmaxz)));
    bminx -= px; bminy -= py; bminz -= pz;
    bmaxx += px; bmaxy += py; bmaxz += pz;
}
SPB_HD void grow_boxes4(f3 o, v4f &minx, v4f &miny, v4f &minz, v4f &maxx, v4f &maxy, v4f &maxz)
{
    grow_box(o, minx.x, miny.x, minz.x, maxx.x, maxy.x, maxz.x);
    grow_box(o, minx.y, miny.y, minz.y, maxx.y, maxy.y, maxz.y);
    grow_box(o, minx.z, miny.z, minz.z, maxx.z, maxy.z, maxz.z);
    grow_box(o, minx.w, miny.w, minz.w, maxx.w, maxy.w, maxz.w);
}

// Moller-Trumbore (ray_intersection.cpp:156-190).  Returns true when the reference would set
// result.t; the caller applies t > 0 (sp_scene.cpp:189).
SPB_HD bool ray_triangle_mt(f3 o, f3 d, f3 a, f3 b, f3 c, float &t, float &u, float &v)
{
    f3 T = sub3(o, a);
    f3 e1 = sub3(b, a);
    f3 e2 = sub3(c, a);
    f3 p = cross3(d, e2);
    f3 q = cross3(T, e1);
    f3 n = cross3(e1, e2);
    float mx = dot3(q, e2), my = dot3(p, T), mz = dot3(q, d);
    float det = 1.0f / dot3(p, e1);
    t = det * mx;
    u = det * my;
    v = det * mz;
    float w = 1.0f - u - v;
    float f = dot3(d, n);
    return (u >= 0.0f && u <= 1.0f && v >= 0.0f && v <= 1.0f && w >= 0.0f && w <= 1.0f &&
            f < 0.0f);
}

// Watertight ray / triangle test (Woop, Benthin, Wald 2013), the north star's option; NEVER the
// parity default: it decides edge cases differently from the reference's Moller-Trumbore by design.
// The triangle is translated to the ray origin and sheared so that the ray runs along +z of a
// permuted frame; the three scaled barycentrics are 2D edge functions of the SAME shared vertices for
// neighbouring triangles, so a ray through a shared edge or vertex cannot fall between them; a zero
// edge function is re-evaluated in double precision.  Same interface and the same acceptance rules
// as ray_triangle_mt(): front faces only (d . (e1 x e2) < 0, ray_intersection.cpp:183-188), (u, v)
// the weights of b and c, the caller applies t > 0.
SPB_HD_NOINLINE bool ray_triangle_wt(f3 o, f3 d, f3 a, f3 b, f3 c, float &t, float &u, float &v)
{
    const float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    int kz = ax > ay ? (ax > az ? 0 : 2) : (ay > az ? 1 : 2);
    int kx = kz == 2 ? 0 : kz + 1, ky = kx == 2 ? 0 : kx + 1;
    const float dv[3] = {d.x, d.y, d.z};
    if (dv[kz] < 0.0f) { int s = kx; kx = ky; ky = s; }
    const float Sx = dv[kx] / dv[kz], Sy = dv[ky] / dv[kz], Sz = 1.0f / dv[kz];
    const float A[3] = {a.x - o.x, a.y - o.y, a.z - o.z}, B[3] = {b.x - o.x, b.y - o.y, b.z - o.z},
                Cv[3] = {c.x - o.x, c.y - o.y, c.z - o.z};
    const float Ax = A[kx] - Sx * A[kz], Ay = A[ky] - Sy * A[kz];
    const float Bx = B[kx] - Sx * B[kz], By = B[ky] - Sy * B[kz];
    const float Cx = Cv[kx] - Sx * Cv[kz], Cy = Cv[ky] - Sy * Cv[kz];
    float U = Cx * By - Cy * Bx, V = Ax * Cy - Ay * Cx, W = Bx * Ay - By * Ax;
    if (U == 0.0f || V == 0.0f || W == 0.0f)
    {
        U = (float)((double)Cx * (double)By - (double)Cy * (double)Bx);
        V = (float)((double)Ax * (double)Cy - (double)Ay * (double)Cx);
        W = (float)((double)Bx * (double)Ay - (double)By * (double)Ax);
    }
    t = -1.0f;
    u = v = 0.0f;
    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
    const float det = U + V + W;
    if (det == 0.0f) return false;
    // front faces only, like the reference
    f3 n = cross3(sub3(b, a), sub3(c, a));
    if (!(dot3(d, n) < 0.0f)) return false;
    const float T = U * (Sz * A[kz]) + V * (Sz * B[kz]) + W * (Sz * Cv[kz]);
    const float rcp = 1.0f / det;
    t = T * rcp;
    u = V * rcp;
    v = W * rcp;
    return true;
}

// the triangle test of the scene (sp_b200_Params::triangleTest; uniform over a launch)
SPB_HD bool ray_triangle(uint32_t test, f3 o, f3 d, f3 a, f3 b, f3 c, float &t, float &u, float &v)
{
    return test == 0 ? ray_triangle_mt(o, d, a, b, c, t, u, v) : ray_triangle_wt(o, d, a, b, c, t, u, v);
}

// ------------------------------------------------------------------------------------------
// traversal

struct Hit
{
    float t;        // world t (sp_scene.cpp:302) or -1
    int32_t object; // object index or -1
    uint32_t slot;  // triangle slot in DScene::tris
    float u, v;     // barycentrics of the winning triangle
    f3 localOrigin, localDirection; // the object-space ray the winner was found with
    float localT;
};

SPB_HD void sort2(float &ta, uint32_t &ra, float &tb, uint32_t &rb)
{
    if (tb < ta)
    {
        float t = ta; ta = tb; tb = t;
        uint32_t r = ra; ra = rb; rb = r;
    }
}

// One BVH over `nodes` starting at node `root`.  LEAF(slot, tnear) is invoked for every
// primitive child whose own box passes the slab predicate; it returns the updated cull
// distance (or the old one).  With CULL the near child is visited first and subtrees whose
// entry distance exceeds the cull distance are skipped; without it every intersected leaf is
// visited, like bvh_IntersectRay (bvh.cpp:203-311).
template <bool CULL, bool EXACT, class LeafFn>
SPB_HD void traverse(const v4f *nodes, uint32_t root, f3 o, f3 inv, float tcull, uint32_t *stack,
                     float *stackT, int stackBase, int stackLimit, Counters *counters,
                     LeafFn &leaf, uint32_t watertight = 0)
{
    int sp = stackBase;
    uint32_t node = root;
    const float inf = u2f(0x7F800000u);
    for (;;)
    {
        const v4f *n = nodes + (size_t)node * 8;
        v4f minx = ld4(n + 0), miny = ld4(n + 1), minz = ld4(n + 2);
        v4f maxx = ld4(n + 3), maxy = ld4(n + 4), maxz = ld4(n + 5);
        v4u refs = ld4u((const v4u *)(n + 6));
        if (counters) counters->nodeVisits++;
        if (watertight) grow_boxes4(o, minx, miny, minz, maxx, maxy, maxz);

        float tn0, tn1, tn2, tn3;
        bool h0, h1, h2, h3;
        if (EXACT)
        {
            h0 = slab_exact(minx.x, miny.x, minz.x, maxx.x, maxy.x, maxz.x, o, inv, tn0);
            h1 = slab_exact(minx.y, miny.y, minz.y, maxx.y, maxy.y, maxz.y, o, inv, tn1);
            h2 = slab_exact(minx.z, miny.z, minz.z, maxx.z, maxy.z, maxz.z, o, inv, tn2);
            h3 = slab_exact(minx.w, miny.w, minz.w, maxx.w, maxy.w, maxz.w, o, inv, tn3);
        }
        else
        {
            h0 = slab_fast(minx.x, miny.x, minz.x, maxx.x, maxy.x, maxz.x, o, inv, tn0);
            h1 = slab_fast(minx.y, miny.y, minz.y, maxx.y, maxy.y, maxz.y, o, inv, tn1);
            h2 = slab_fast(minx.z, miny.z, minz.z, maxx.z, maxy.z, maxz.z, o, inv, tn2);
            h3 = slab_fast(minx.w, miny.w, minz.w, maxx.w, maxy.w, maxz.w, o, inv, tn3);
        }
        h0 = h0 && refs.x != SPB_REF_EMPTY;
        h1 = h1 && refs.y != SPB_REF_EMPTY;
        h2 = h2 && refs.z != SPB_REF_EMPTY;
        h3 = h3 && refs.w != SPB_REF_EMPTY;

        // primitives first (in child order), so the cull distance shrinks before descending
        if (h0 && (refs.x & SPB_REF_LEAF)) { if (!CULL || tn0 <= tcull) tcull = leaf(refs.x & ~SPB_REF_LEAF, tcull); h0 = false; }
        if (h1 && (refs.y & SPB_REF_LEAF)) { if (!CULL || tn1 <= tcull) tcull = leaf(refs.y & ~SPB_REF_LEAF, tcull); h1 = false; }
        if (h2 && (refs.z & SPB_REF_LEAF)) { if (!CULL || tn2 <= tcull) tcull = leaf(refs.z & ~SPB_REF_LEAF, tcull); h2 = false; }
        if (h3 && (refs.w & SPB_REF_LEAF)) { if (!CULL || tn3 <= tcull) tcull = leaf(refs.w & ~SPB_REF_LEAF, tcull); h3 = false; }

        if (CULL)
        {
            h0 = h0 && tn0 <= tcull;
            h1 = h1 && tn1 <= tcull;
            h2 = h2 && tn2 <= tcull;
            h3 = h3 && tn3 <= tcull;
        }
        float k0 = h0 ? tn0 : inf, k1 = h1 ? tn1 : inf, k2 = h2 ? tn2 : inf, k3 = h3 ? tn3 : inf;
        uint32_t r0 = h0 ? refs.x : SPB_REF_EMPTY, r1 = h1 ? refs.y : SPB_REF_EMPTY;
        uint32_t r2 = h2 ? refs.z : SPB_REF_EMPTY, r3 = h3 ? refs.w : SPB_REF_EMPTY;
        if (CULL)
        {
            // 5-exchange network, ascending by entry distance; invalid lanes (inf) sink
            sort2(k0, r0, k1, r1);
            sort2(k2, r2, k3, r3);
            sort2(k0, r0, k2, r2);
            sort2(k1, r1, k3, r3);
            sort2(k1, r1, k2, r2);
        }
        // push far to near; continue with the nearest
        // (the builder bounds the depth so the limit is never reached; an entry that would
        // overflow is dropped rather than written out of bounds)
        if (r3 != SPB_REF_EMPTY && sp < stackLimit) { stack[sp] = r3; if (CULL) stackT[sp] = k3; sp++; }
        if (r2 != SPB_REF_EMPTY && sp < stackLimit) { stack[sp] = r2; if (CULL) stackT[sp] = k2; sp++; }
        if (r1 != SPB_REF_EMPTY && sp < stackLimit) { stack[sp] = r1; if (CULL) stackT[sp] = k1; sp++; }
        if (r0 != SPB_REF_EMPTY)
        {
            node = r0;
            continue;
        }
        // pop
        bool found = false;
        while (sp > stackBase)
        {
            sp--;
            if (!CULL || stackT[sp] <= tcull)
            {
                node = stack[sp];
                found = true;
                break;
            }
        }
        if (!found) break;
    }
}

// Slack applied to the running closest distance before it is used to skip subtrees: the slab
// entry distance of a triangle's own box and its Moller-Trumbore t are different roundings of
// the same quantity, so a strict comparison could drop an equal-distance winner.
#define SPB_CULL_SLACK 1.0009765625f /* 1 + 2^-10 */

// A WORLD distance (sp_scene.cpp:296-302: t = Dot(M * localHit - origin, dir)) comes out of a
// transform and a subtraction of coordinates, so its error is absolute -- a few ulps of the
// coordinates involved -- not relative to t.  For a short ray far from the scene origin that can
// exceed the relative slack, so every bound derived from a world t is first padded by 2^-16 of
// (|origin|_1 + |t|) (>= 100x the worst rounding of the transform chain).  Only object entry and
// exit pay for it.
SPB_HD float cull_pad(f3 worldOrigin, float t)
{
    return 1.52587890625e-05f * (fabsf(worldOrigin.x) + fabsf(worldOrigin.y) + fabsf(worldOrigin.z) + fabsf(t));
}

// sp_RayIntersectMesh (sp_scene.cpp:127-227) on one object-space ray.
// WT: whether sp_b200_Params::triangleTest may be honoured here (false in the wavefront kernels, which
// run the reference's test only and must not carry the option's code at all).
template <bool CULL, bool EXACT, bool WT = true>
SPB_HD void intersect_mesh(const DScene &S, uint32_t meshRoot, f3 o, f3 d, float tcull,
                           uint32_t *stack, float *stackT, int stackBase, Counters *counters,
                           float &bestT, uint32_t &bestSlot, float &bestU, float &bestV)
{
    bestT = -1.0f;
    bestSlot = 0;
    bestU = bestV = 0.0f;
    if (meshRoot == SPB_REF_EMPTY) return;
    f3 inv = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z); // Inverse(), math_lib.h:911-915
    auto leaf = [&](uint32_t slot, float cull) -> float {
        const v4f *tp = S.tris + (size_t)slot * 3;
        v4f a = ld4(tp + 0), b = ld4(tp + 1), c = ld4(tp + 2);
        if (counters) counters->triangleTests++;
        float t, u, v;
        if (ray_triangle(WT ? S.triangleTest : 0u, o, d, mk3(a.x, a.y, a.z), mk3(b.x, b.y, b.z), mk3(c.x, c.y, c.z), t, u, v))
        {
            if (t > 0.0f)
            {
                if (t < bestT || bestT < 0.0f)
                {
                    bestT = t;
                    bestSlot = slot;
                    bestU = u;
                    bestV = v;
                    float c2 = t * SPB_CULL_SLACK;
                    if (c2 < cull) cull = c2;
                }
            }
        }
        return cull;
    };
    traverse<CULL, EXACT>(S.nodes, meshRoot, o, inv, tcull, stack, stackT, stackBase, SPB_STACK_SIZE, counters, leaf,
                          WT ? S.triangleTest : 0u);
}

SPB_HD m4 load_m4(const v4f *p)
{
    m4 m;
    m.c[0] = ld4(p + 0);
    m.c[1] = ld4(p + 1);
    m.c[2] = ld4(p + 2);
    m.c[3] = ld4(p + 3);
    return m;
}

SPB_HD bool any_nonfinite_inv(f3 d)
{
    // 1/d is +-inf when |d| is zero or a tiny denormal; only then can (b-o)*inv be NaN
    const float tiny = 2.938736e-39f; // 1/tiny is still finite (FLT_MAX ~ 3.4e38)
    return !(fabsf(d.x) > tiny && fabsf(d.y) > tiny && fabsf(d.z) > tiny);
}

// sp_RayIntersectScene (sp_scene.cpp:229-339): closest hit over all objects whose world AABB the
// ray passes.  The winner's surface attributes are resolved afterwards by resolve_hit().
template <bool CULL, bool WT = true>
SPB_HD Hit intersect_scene(const DScene &S, f3 o, f3 d, uint32_t *stack, float *stackT,
                           Counters *counters)
{
    Hit best;
    best.t = -1.0f;
    best.object = -1;
    best.slot = 0;
    best.u = best.v = 0.0f;
    best.localOrigin = best.localDirection = mk3(0, 0, 0);
    best.localT = -1.0f;
    if (S.tlasRoot == SPB_REF_EMPTY) return best;

    f3 inv = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    const float inf = u2f(0x7F800000u);
    const bool worldExact = any_nonfinite_inv(d);

    auto objectLeaf = [&](uint32_t objectIndex, float cull) -> float {
        v4u info = ld4u(S.objInfo + objectIndex);
        m4 invModel = load_m4(S.objInv + (size_t)objectIndex * 4);
        if (counters) counters->objectTests++;
        // sp_scene.cpp:274-276
        f3 lo = xform(invModel, o, 1.0f);
        float scaleLen;
        f3 ld = normalize3(xform(invModel, d, 0.0f), &scaleLen);

        // distance bound in object space: |M^-1 d| maps world t to local t exactly (affine)
        float localCull = inf;
        if (CULL && best.t >= 0.0f) localCull = (best.t + cull_pad(o, best.t)) * scaleLen * SPB_CULL_SLACK;

        float lt, lu, lv;
        uint32_t lslot;
        // the object traversal continues on the same stack above the entries of the TLAS
        int base = SPB_STACK_SIZE / 3;
        if (any_nonfinite_inv(ld))
            intersect_mesh<CULL, true, WT>(S, info.x, lo, ld, localCull, stack, stackT, base, counters, lt, lslot, lu, lv);
        else
            intersect_mesh<CULL, false, WT>(S, info.x, lo, ld, localCull, stack, stackT, base, counters, lt, lslot, lu, lv);

        if (lt >= 0.0f)
        {
            // sp_scene.cpp:296-306
            m4 model = load_m4(S.objModel + (size_t)objectIndex * 4);
            f3 localHit = add3(lo, mul3(ld, lt));
            f3 worldHit = xform(model, localHit, 1.0f);
            float t = dot3(sub3(worldHit, o), d);
            if (t < best.t || best.t < 0.0f)
            {
                best.t = t;
                best.object = (int32_t)objectIndex;
                best.slot = lslot;
                best.u = lu;
                best.v = lv;
                best.localOrigin = lo;
                best.localDirection = ld;
                best.localT = lt;
                float c2 = (t + cull_pad(o, t)) * SPB_CULL_SLACK;
                if (t > 0.0f && c2 < cull) cull = c2;
            }
        }
        return cull;
    };

    const uint32_t watertight = WT ? S.triangleTest : 0u;
    if (worldExact)
        traverse<CULL, true>(S.nodes, S.tlasRoot, o, inv, inf, stack, stackT, 0, SPB_STACK_SIZE / 3, counters, objectLeaf, watertight);
    else
        traverse<CULL, false>(S.nodes, S.tlasRoot, o, inv, inf, stack, stackT, 0, SPB_STACK_SIZE / 3, counters, objectLeaf, watertight);
    return best;
}

// ------------------------------------------------------------------------------------------
// Resumable traversal: the same closest-hit query as intersect_scene(), written as a state
// machine so that a warp can advance 32 independent rays one step at a time, retire the
// finished ones and refill their lanes (spb_wavefront.cu).  A step is either a NODE step (fetch
// a 4-wide node, slab-test its children, push the children that pass, nearest on top) or a LEAF
// step (inside an object: Moller-Trumbore on one triangle whose own box passed; in the TLAS:
// enter one object).  Leaves go through the stack like nodes, so a warp can run the two kinds of
// step separately with most of its lanes active; entries are visited near to far and skipped when
// their entry distance exceeds the closest hit so far.  The closest hit is the minimum over the
// same set of (own box passes) && (Moller-Trumbore accepts) && (t > 0) triangles as
// intersect_scene(); only the winner among exactly equal t can differ (order of visits).
// Only the fast slab form is used here: a ray (world or object space) whose reciprocal direction
// is not finite sets `slow` and the caller finishes it with intersect_scene().
#define SPB_NODE_DONE 0xFFFFFFFEu
// the walk inside an object has ended: trav_exit() must run before anything else (kept as a
// separate step so that a warp can run it for several lanes at once)
#define SPB_NODE_EXIT 0xFFFFFFFDu

// One stack entry: reference and entry distance together (one 8-byte local load per pop).
struct SPB_ALIGN8 TravEntry { uint32_t ref; float tnear; };
SPB_HD TravEntry load_entry(const TravEntry *p)
{
#if defined(__CUDA_ARCH__)
    // both words at once, whether or not the entry turns out to be culled (volatile keeps the
    // compiler from splitting it into a load of tnear and a dependent load of ref)
    unsigned long long raw = *reinterpret_cast<const volatile unsigned long long *>(p);
    TravEntry e;
    e.ref = (uint32_t)raw;
    e.tnear = __uint_as_float((uint32_t)(raw >> 32));
    return e;
#else
    return *p;
#endif
}

// Where a lane's stack lives.  The traversal functions below take any STK with stack_put / stack_get: a plain
// array (host builds, per-pixel and query kernels), or the trace kernel's hybrid (spb_wavefront.cu: the first
// entries -- all a walk needs 99.9 % of the time -- in shared memory, one column per thread, so a push or a pop
// is two conflict-free wavefronts however far the lanes' stack pointers have drifted apart; a local-memory
// access costs two per distinct stack pointer in the warp, on the L1 the node fetches need).
SPB_HD void stack_put(TravEntry *stack, int sp, TravEntry e) { stack[sp] = e; }
SPB_HD TravEntry stack_get(const TravEntry *stack, int sp) { return load_entry(stack + sp); }

// Hot state: what every NODE / LEAF step touches.
struct Trav
{
    f3 o, d, inv;     // ray in the current space (world in the TLAS, object space inside an object)
    float tcull;      // cull distance in the current space
    uint32_t cur;     // entry to process next: node index, SPB_REF_LEAF | slot, SPB_NODE_EXIT, SPB_NODE_DONE
    int sp;           // stack pointer
    int blasBase;     // -1 in the TLAS, else the stack pointer at object entry
    float lT;         // closest hit inside the current object (object space)
    uint32_t lSlot;
};

// Cold state: touched only when an object is entered or left and when the ray is retired (the
// trace kernel keeps it in shared memory, off the register file).
struct TravCold
{
    float worldCull;  // TLAS cull distance, parked while inside an object
    uint32_t object;  // object being traversed
    float bT;         // closest hit so far (world t)
    uint32_t bSlot;
    int32_t bObject;
    uint32_t slow;
};

SPB_HD bool trav_is_node(const Trav &st) { return (st.cur & SPB_REF_LEAF) == 0; } // DONE / EXIT have the bit set
SPB_HD bool trav_is_walking(const Trav &st) { return st.cur < SPB_NODE_EXIT; }
SPB_HD bool trav_is_walking_ref(uint32_t cur) { return cur < SPB_NODE_EXIT; }

SPB_HD void trav_begin(const DScene &S, f3 o, f3 d, Trav &st, TravCold &c)
{
    st.o = o;
    st.d = d;
    st.inv = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    st.tcull = u2f(0x7F800000u);
    st.sp = 0;
    st.blasBase = -1;
    st.lT = -1.0f; st.lSlot = 0;
    c.worldCull = st.tcull;
    c.object = 0;
    c.bT = -1.0f; c.bSlot = 0; c.bObject = -1;
    c.slow = 0;
    st.cur = S.tlasRoot == SPB_REF_EMPTY ? SPB_NODE_DONE : S.tlasRoot;
    if (st.cur != SPB_NODE_DONE && any_nonfinite_inv(d))
    {
        c.slow = 1;
        st.cur = SPB_NODE_DONE;
    }
}

// `ray` points at the world ray record (two float4: origin, direction); it is only read when an
// object is entered or left, so the world ray does not occupy registers during the walk.
SPB_HD void trav_world_ray(const v4f *ray, f3 &wo, f3 &wd)
{
    v4f a = ray[0], b = ray[1];
    wo = mk3(a.x, a.y, a.z);
    wd = mk3(b.x, b.y, b.z);
}

// Pops until there is an entry to process (st.cur), the walk inside an object ends
// (SPB_NODE_EXIT) or the query ends (SPB_NODE_DONE).
// SINGLE (here and in the steps below): the walk is known to be inside the one object of a single-object
// scene from its first step to its last (trav_begin_single) -- the stack base is 0, an empty stack is the
// end of the object, and a leaf is always a triangle; the trace kernel is compiled once for such scenes.
template <bool CULL, bool SINGLE = false, class STK>
SPB_HD void trav_pop(Trav &st, STK stack)
{
    for (;;)
    {
        int base = SINGLE ? 0 : (st.blasBase >= 0 ? st.blasBase : 0);
        if (st.sp == base)
        {
            st.cur = (SINGLE || st.blasBase >= 0) ? SPB_NODE_EXIT : SPB_NODE_DONE;
            return;
        }
        st.sp--;
        TravEntry e = stack_get(stack, st.sp);
        if (CULL && !(e.tnear <= st.tcull)) continue;
        st.cur = e.ref;
        return;
    }
}

// EXIT step (st.cur == SPB_NODE_EXIT): leave the object -- sp_scene.cpp:296-322 on its closest
// hit -- and go on with the TLAS entries below.
template <bool CULL, class STK>
SPB_HD void trav_exit(const DScene &S, Trav &st, TravCold &c, const v4f *ray, STK stack)
{
    f3 wo, wd;
    trav_world_ray(ray, wo, wd);
    float worldCull = c.worldCull;
    if (st.lT >= 0.0f)
    {
        m4 model = load_m4(S.objModel + (size_t)c.object * 4);
        f3 localHit = add3(st.o, mul3(st.d, st.lT));
        f3 worldHit = xform(model, localHit, 1.0f);
        float t = dot3(sub3(worldHit, wo), wd);
        float bT = c.bT;
        if (t < bT || bT < 0.0f)
        {
            c.bT = t;
            c.bObject = (int32_t)c.object;
            c.bSlot = st.lSlot;
            float c2 = (t + cull_pad(wo, t)) * SPB_CULL_SLACK;
            if (t > 0.0f && c2 < worldCull) worldCull = c2;
        }
    }
    st.o = wo;
    st.d = wd;
    st.inv = mk3(1.0f / wd.x, 1.0f / wd.y, 1.0f / wd.z);
    st.tcull = worldCull;
    st.blasBase = -1;
    trav_pop<CULL>(st, stack);
}

// The exit arithmetic alone (sp_scene.cpp:296-322): the object's closest hit carried to world space.
SPB_HD void trav_leave(const DScene &S, const Trav &st, TravCold &c, const v4f *ray)
{
    if (!(st.lT >= 0.0f)) return;
    f3 wo, wd;
    trav_world_ray(ray, wo, wd);
    m4 model = load_m4(S.objModel + (size_t)c.object * 4);
    f3 localHit = add3(st.o, mul3(st.d, st.lT));
    f3 worldHit = xform(model, localHit, 1.0f);
    float t = dot3(sub3(worldHit, wo), wd);
    float bT = c.bT;
    if (t < bT || bT < 0.0f)
    {
        c.bT = t;
        c.bObject = (int32_t)c.object;
        c.bSlot = st.lSlot;
    }
}

// Single-object scenes: what the walk does before it reaches the mesh tree -- the TLAS root's one
// child is the object: the exact test of its box, then the object entry of trav_leaf() -- done at
// once when the ray starts, with every lane of a refilling warp busy; the walk then ends with
// SPB_NODE_EXIT, which the caller treats as finished (trav_leave() at retire).
template <bool CULL>
SPB_HD void trav_begin_single(const DScene &S, f3 o, f3 d, Trav &st, TravCold &c, Counters *counters)
{
    trav_begin(S, o, d, st, c);
    if (st.cur == SPB_NODE_DONE) return;
    st.cur = SPB_NODE_DONE;
    if (counters) counters->nodeVisits++;
    v4f bmin = ld4(S.objBox), bmax = ld4(S.objBox + 1);
    float tn;
    if (!slab_fast(bmin.x, bmin.y, bmin.z, bmax.x, bmax.y, bmax.z, st.o, st.inv, tn)) return;
    v4u info = ld4u(S.objInfo);
    if (counters) counters->objectTests++;
    if (info.x == SPB_REF_EMPTY) return;
    m4 invModel = load_m4(S.objInv);
    f3 lo = xform(invModel, o, 1.0f);
    f3 ld = normalize3(xform(invModel, d, 0.0f));
    if (any_nonfinite_inv(ld))
    {
        c.slow = 1;
        return;
    }
    c.worldCull = st.tcull;
    st.o = lo;
    st.d = ld;
    st.inv = mk3(1.0f / ld.x, 1.0f / ld.y, 1.0f / ld.z);
    c.object = 0;
    st.blasBase = 0;
    st.lT = -1.0f;
    st.lSlot = 0;
    st.cur = info.x;
}

// No bound check: flatten_scene() refuses scenes whose worst-case stack use (three entries per
// level of the TLAS plus three per level of the deepest mesh tree) exceeds SPB_STACK_SIZE.
template <class STK>
SPB_HD void trav_push(Trav &st, STK stack, uint32_t ref, float tnear)
{
    TravEntry e;
    e.ref = ref;
    e.tnear = tnear;
    stack_put(stack, st.sp, e);
    st.sp++;
}

// compare-exchange on (key, ref) with min/max on the keys (keys are never NaN)
SPB_HD void sortx(float &ka, uint32_t &ra, float &kb, uint32_t &rb)
{
    bool swap = kb < ka;
    float lo = fminf(ka, kb), hi = fmaxf(ka, kb);
    uint32_t rlo = swap ? rb : ra, rhi = swap ? ra : rb;
    ka = lo; kb = hi; ra = rlo; rb = rhi;
}

// The 128 bytes of a node as the NODE step reads them.
struct NodeData { v4f minx, miny, minz, maxx, maxy, maxz, refsf, meta; };
// The fetch of a NODE step alone, where `take` is set: the trace kernel issues it as soon as a lane's next
// entry is known to be a node, ahead of the warp's vote on the kind of the next step, so that the L1 round
// trip runs under the vote instead of in front of the first slab test.
SPB_HD void trav_node_fetch(const DScene &S, bool take, uint32_t cur, NodeData &nd)
{
    // (every lane loads: a predicated load would leave its registers "maybe written", which keeps them alive
    // around the whole loop; lanes without a node read node 0, one line for all of them)
    const v4f *n = S.nodes + (size_t)(take ? cur : 0u) * 8;
    ld8(n + 0, nd.minx, nd.miny);
    ld8(n + 2, nd.minz, nd.maxx);
    ld8(n + 4, nd.maxy, nd.maxz);
    ld8(n + 6, nd.refsf, nd.meta);
}
// NODE step on fetched data: st.cur is the node index nd was fetched for.
template <bool CULL, bool SINGLE = false, class STK>
SPB_HD void trav_node_apply(const DScene &S, Trav &st, STK stack, Counters *counters, const NodeData &nd)
{
    const float inf = u2f(0x7F800000u);
    const v4f minx = nd.minx, miny = nd.miny, minz = nd.minz, maxx = nd.maxx, maxy = nd.maxy, maxz = nd.maxz, refsf = nd.refsf;
    v4u refs;
    refs.x = f2u(refsf.x); refs.y = f2u(refsf.y); refs.z = f2u(refsf.z); refs.w = f2u(refsf.w);
    if (counters) counters->nodeVisits++;

    float tn0, tn1, tn2, tn3;
    bool h0 = slab_fast(minx.x, miny.x, minz.x, maxx.x, maxy.x, maxz.x, st.o, st.inv, tn0);
    bool h1 = slab_fast(minx.y, miny.y, minz.y, maxx.y, maxy.y, maxz.y, st.o, st.inv, tn1);
    bool h2 = slab_fast(minx.z, miny.z, minz.z, maxx.z, maxy.z, maxz.z, st.o, st.inv, tn2);
    bool h3 = slab_fast(minx.w, miny.w, minz.w, maxx.w, maxy.w, maxz.w, st.o, st.inv, tn3);
    // an empty child slot carries a NaN box (spb_bvh.cpp collapse): tfar is NaN, the test fails
    if (CULL)
    {
        h0 = h0 && tn0 <= st.tcull;
        h1 = h1 && tn1 <= st.tcull;
        h2 = h2 && tn2 <= st.tcull;
        h3 = h3 && tn3 <= st.tcull;
    }
    // a child that is not entered gets the key +inf (an entry distance of +inf cannot lead to a
    // hit either); validity travels with the key, the refs are never rewritten
    float k0 = h0 ? tn0 : inf, k1 = h1 ? tn1 : inf, k2 = h2 ? tn2 : inf, k3 = h3 ? tn3 : inf;
    uint32_t r0 = refs.x, r1 = refs.y, r2 = refs.z, r3 = refs.w;
    if (CULL)
    {
        // 5-exchange network, ascending by entry distance; +inf sinks
        sortx(k0, r0, k1, r1);
        sortx(k2, r2, k3, r3);
        sortx(k0, r0, k2, r2);
#if !defined(SPB_SORT_NEAREST_ONLY)
        sortx(k1, r1, k3, r3);
        sortx(k1, r1, k2, r2);
#endif
    }
    else
    {
        // no distance order: keep child order, valid entries first
        if (!(k0 < inf)) { k0 = k1; r0 = r1; k1 = inf; }
        if (!(k1 < inf)) { k1 = k2; r1 = r2; k2 = inf; }
        if (!(k2 < inf)) { k2 = k3; r2 = r3; k3 = inf; }
        if (!(k0 < inf)) { k0 = k1; r0 = r1; k1 = inf; }
        if (!(k1 < inf)) { k1 = k2; r1 = r2; k2 = inf; }
        if (!(k0 < inf)) { k0 = k1; r0 = r1; k1 = inf; }
    }
#if defined(SPB_SORT_NEAREST_ONLY)
    // (A/B knob) only the nearest child is exact; k2 is the nearer of the two pair minima
    if (k3 < inf) trav_push(st, stack, r3, k3);
    if (k1 < inf) trav_push(st, stack, r1, k1);
    if (k2 < inf) trav_push(st, stack, r2, k2);
#else
    if (k3 < inf) trav_push(st, stack, r3, k3);
    if (k2 < inf) trav_push(st, stack, r2, k2);
    if (k1 < inf) trav_push(st, stack, r1, k1);
#endif
    if (k0 < inf)
    {
        st.cur = r0;
        return;
    }
    trav_pop<CULL, SINGLE>(st, stack);
}
// NODE step: st.cur is a node index.
template <bool CULL, bool SINGLE = false, class STK>
SPB_HD void trav_node(const DScene &S, Trav &st, STK stack, Counters *counters)
{
    NodeData nd;
    const v4f *n = S.nodes + (size_t)st.cur * 8;
    ld8(n + 0, nd.minx, nd.miny);
    ld8(n + 2, nd.minz, nd.maxx);
    ld8(n + 4, nd.maxy, nd.maxz);
    ld8(n + 6, nd.refsf, nd.meta);
    trav_node_apply<CULL, SINGLE>(S, st, stack, counters, nd);
}

// LEAF step: st.cur is SPB_REF_LEAF | slot.
template <bool CULL, bool SINGLE = false, class STK>
SPB_HD void trav_leaf(const DScene &S, Trav &st, TravCold &c, const v4f *ray, STK stack,
                      Counters *counters)
{
    const float inf = u2f(0x7F800000u);
    uint32_t index = st.cur & ~SPB_REF_LEAF;
    if (SINGLE || st.blasBase >= 0)
    {
        // a triangle whose own box the ray passes (sp_scene.cpp:161-196)
        const v4f *tp = S.tris + (size_t)index * 3;
        v4f a = ld4(tp + 0), b = ld4(tp + 1), cc = ld4(tp + 2);
        if (counters) counters->triangleTests++;
        float t, u, v;
        if (ray_triangle_mt(st.o, st.d, mk3(a.x, a.y, a.z), mk3(b.x, b.y, b.z), mk3(cc.x, cc.y, cc.z), t, u, v))
        {
            if (t > 0.0f && (t < st.lT || st.lT < 0.0f))
            {
                st.lT = t;
                st.lSlot = index;
                float c2 = t * SPB_CULL_SLACK;
                if (c2 < st.tcull) st.tcull = c2;
            }
        }
    }
    else
    {
        // enter an object (sp_scene.cpp:274-276)
        v4u info = ld4u(S.objInfo + index);
        if (counters) counters->objectTests++;
        if (info.x != SPB_REF_EMPTY)
        {
            m4 invModel = load_m4(S.objInv + (size_t)index * 4);
            f3 wo, wd;
            trav_world_ray(ray, wo, wd);
            f3 lo = xform(invModel, wo, 1.0f);
            float scaleLen;
            f3 ld = normalize3(xform(invModel, wd, 0.0f), &scaleLen);
            if (any_nonfinite_inv(ld))
            {
                c.slow = 1;
                st.cur = SPB_NODE_DONE;
                return;
            }
            c.worldCull = st.tcull;
            st.tcull = inf;
            float bT = c.bT;
            if (CULL && bT >= 0.0f) st.tcull = (bT + cull_pad(wo, bT)) * scaleLen * SPB_CULL_SLACK;
            st.o = lo;
            st.d = ld;
            st.inv = mk3(1.0f / ld.x, 1.0f / ld.y, 1.0f / ld.z);
            c.object = index;
            st.blasBase = st.sp;
            st.lT = -1.0f;
            st.lSlot = 0;
            st.cur = info.x;
            return;
        }
    }
    trav_pop<CULL, SINGLE>(st, stack);
}

// Barycentrics of the winning triangle, recomputed from the same inputs with the same code the
// traversal ran (so the same bits): the traversal itself carries only (t, slot, object).
// `watertight`: the triangle test the hit was found with (a compile-time 0 in the wavefront kernels,
// which only run the reference's test; DScene::triangleTest in the per-pixel and query kernels).
SPB_HD void hit_barycentrics(const DScene &S, f3 o, f3 d, Hit &hit, uint32_t watertight = 0)
{
    m4 invModel = load_m4(S.objInv + (size_t)hit.object * 4);
    f3 lo = xform(invModel, o, 1.0f);
    f3 ld = normalize3(xform(invModel, d, 0.0f));
    const v4f *tp = S.tris + (size_t)hit.slot * 3;
    v4f a = ld4(tp + 0), b = ld4(tp + 1), c = ld4(tp + 2);
    float t;
    ray_triangle(watertight, lo, ld, mk3(a.x, a.y, a.z), mk3(b.x, b.y, b.z), mk3(c.x, c.y, c.z), t, hit.u, hit.v);
}

SPB_HD Hit trav_result(const TravCold &c)
{
    Hit h;
    h.t = c.bT;
    h.object = c.bObject;
    h.slot = c.bSlot;
    h.u = h.v = 0.0f;
    h.localOrigin = h.localDirection = mk3(0, 0, 0);
    h.localT = -1.0f;
    return h;
}

// The state machine run to completion for one ray (what each lane of k_trace does, unrolled in
// time, single-object entry at the start included); hostsim uses it to check the machine against
// intersect_scene() and the oracle.
template <bool CULL>
SPB_HD Hit intersect_scene_stepped(const DScene &S, f3 o, f3 d, uint32_t *stack, float *stackT,
                                   Counters *counters)
{
    Trav st;
    TravCold cold;
    TravEntry entries[SPB_STACK_SIZE];
    v4f ray[2];
    ray[0].x = o.x; ray[0].y = o.y; ray[0].z = o.z; ray[0].w = 0.0f;
    ray[1].x = d.x; ray[1].y = d.y; ray[1].z = d.z; ray[1].w = 0.0f;
    const bool single = S.objectCount == 1;
    if (single) trav_begin_single<CULL>(S, o, d, st, cold, counters);
    else trav_begin(S, o, d, st, cold);
    while (st.cur != SPB_NODE_DONE && !(single && st.cur == SPB_NODE_EXIT))
    {
        if (st.cur == SPB_NODE_EXIT) trav_exit<CULL>(S, st, cold, ray, entries);
        else if (trav_is_node(st)) trav_node<CULL>(S, st, entries, counters);
        else trav_leaf<CULL>(S, st, cold, ray, entries, counters);
    }
    if (single && st.cur == SPB_NODE_EXIT) trav_leave(S, st, cold, ray);
    if (cold.slow) return intersect_scene<CULL, false>(S, o, d, stack, stackT, counters);
    Hit h = trav_result(cold);
    if (h.object >= 0) hit_barycentrics(S, o, d, h);
    return h;
}

// ------------------------------------------------------------------------------------------
// Resumable traversal, second form (round 2; what k_trace runs).  Same query, same result set as
// intersect_scene() -- the minimum over {triangles whose OWN box passes the reference's slab
// predicate, Moller-Trumbore accepts, t > 0} over {objects whose OWN world box passes it} -- but
// inside an object the reference's predicate is evaluated only where the result depends on it:
//
//   * the boxes of a mesh tree (inner boxes, and the first look at a leaf's box inside its
//     parent's node step) get a CONSERVATIVE test that costs half the instructions and none of the
//     min/max pairs: per axis
//         tnear = bmin * p + bmax * q + cn,   tfar = bmin * q + bmax * p + cf      (four FFMA)
//     with (p, q) = (1/d, 0) or (0, 1/d) by the sign of d -- the multiply by zero selects the near
//     and the far plane without a compare -- and cn / cf = -(o -+ delta) / d, the box grown by
//     delta = 2^-21 (|o|_1 + extent of the tree) on every side.  Both forms round (b - o) / d a few
//     ulps of (|b| + |o|) / |d| away from its value (the reference: sub, mul; this one: mul, fma);
//     delta is eight times the sum of the two bounds, so the padded interval of every axis
//     contains the reference's and whatever passes the reference's test passes this one.
//   * a triangle that comes off the stack is tested AGAIN, exactly: its own box is rebuilt from
//     its three vertices (min / max are exact, so these are the builder's bits up to the sign of
//     a zero, which no comparison of the slab test can see) and slab_fast() on it is the
//     reference's test (all reciprocals finite).  Only then Moller-Trumbore runs.
//   * the TLAS is walked with the exact test throughout (slab_fast on the world ray, which stays
//     in the lane's record): its leaves are the objects' own boxes, few nodes are visited per
//     ray, and a world-space set of test constants would have to be parked and restored around
//     every object (measured on the 182-object scene: 10 % of the frame).
//
// The set of leaves that reach Moller-Trumbore / the object entry is therefore exactly the
// reference's; the conservative test only decides which subtrees of a mesh are opened on the way.
// Rays with a reciprocal direction beyond 1e30 (|d| < 1e-30 on some axis, where b * (1/d) could
// overflow) or coordinates beyond 1e8 take the exact non-resumable walk, like rays with a
// non-finite reciprocal before.
//
// State.  Registers ("hot", Trav2): the twelve test constants of the object being walked, the
// cull distance, the current entry, the stack pointer and the object (NONE in the TLAS).
// Everything else lives in a per-lane record reached through a strided view (shared memory in the
// kernel, a plain array on the host): the world ray and its reciprocal, the object-space ray and
// its reciprocal for the exact triangle tests, the closest hit inside the object and overall.
// The stack carries sentinel entries instead of a base pointer: {DONE} at the bottom, {EXIT}
// under the entries of an object.
struct Trav2
{
    f3 p, q, cn, cf;
    float tcull;
    uint32_t cur;
    int sp;
    uint32_t object; // SPB_T2_NO_OBJECT while in the TLAS
};

enum
{
    T2_OX = 0, T2_OY, T2_OZ, T2_DX, T2_DY, T2_DZ, T2_IX, T2_IY, T2_IZ,       // object-space ray, reciprocal
    T2_LT, T2_LSLOT,                                                         // closest hit inside the object
    T2_WOX, T2_WOY, T2_WOZ, T2_WDX, T2_WDY, T2_WDZ, T2_WIX, T2_WIY, T2_WIZ,  // world ray, reciprocal
    T2_WORLDCULL, T2_BT, T2_BSLOT, T2_BOBJECT, T2_SLOW,                      // world cull distance while inside an object; closest hit overall
    T2_WORDS
};
#define SPB_T2_NO_OBJECT 0xFFFFFFFFu

// one lane's record: word k at base[k * STRIDE]
template <int STRIDE>
struct T2View
{
    float *base;
    SPB_HD float &f(int k) const { return base[k * STRIDE]; }
    SPB_HD uint32_t &u(int k) const { return reinterpret_cast<uint32_t *>(base)[k * STRIDE]; }
    SPB_HD f3 f3at(int k) const { return mk3(base[k * STRIDE], base[(k + 1) * STRIDE], base[(k + 2) * STRIDE]); }
    SPB_HD void set3(int k, f3 v) const { base[k * STRIDE] = v.x; base[(k + 1) * STRIDE] = v.y; base[(k + 2) * STRIDE] = v.z; }
};

SPB_HD float fma_rn(float a, float b, float c)
{
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(a, b, c);
#else
    return fmaf(a, b, c);
#endif
}

// |d| >= 1e-30 on every axis and every coordinate below 1e8: b * (1/d) and o * (1/d) stay finite
SPB_HD bool trav2_ray_ok(f3 o, f3 d, float extent)
{
    const float tiny = 1.0e-30f;
    const float big = fabsf(o.x) + fabsf(o.y) + fabsf(o.z) + extent;
    return fabsf(d.x) >= tiny && fabsf(d.y) >= tiny && fabsf(d.z) >= tiny && big < 1.0e8f;
}

// test constants of the ray (o, d) against a tree whose boxes lie within +-extent
SPB_HD void trav2_constants(f3 o, f3 inv, float extent, Trav2 &st)
{
    const float delta = 4.76837158203125e-07f * (fabsf(o.x) + fabsf(o.y) + fabsf(o.z) + extent); // 2^-21
    const float kminx = -(o.x + delta) * inv.x, kmaxx = -(o.x - delta) * inv.x; // pairs with bmin / bmax
    const float kminy = -(o.y + delta) * inv.y, kmaxy = -(o.y - delta) * inv.y;
    const float kminz = -(o.z + delta) * inv.z, kmaxz = -(o.z - delta) * inv.z;
    const bool px = inv.x >= 0.0f, py = inv.y >= 0.0f, pz = inv.z >= 0.0f;
    st.p = mk3(px ? inv.x : 0.0f, py ? inv.y : 0.0f, pz ? inv.z : 0.0f);
    st.q = mk3(px ? 0.0f : inv.x, py ? 0.0f : inv.y, pz ? 0.0f : inv.z);
    st.cn = mk3(px ? kminx : kmaxx, py ? kminy : kmaxy, pz ? kminz : kmaxz);
    st.cf = mk3(px ? kmaxx : kminx, py ? kmaxy : kminy, pz ? kmaxz : kminz);
}

// conservative slab test of one child (see above).  Returns the key the children are ordered by:
// the entry distance if the padded box is entered before `tlimit` (the cull distance), else +inf.
SPB_HD float slab_wide(float bminx, float bminy, float bminz, float bmaxx, float bmaxy, float bmaxz, const Trav2 &st,
                       float tlimit)
{
    const float tnx = fma_rn(bminx, st.p.x, fma_rn(bmaxx, st.q.x, st.cn.x));
    const float tny = fma_rn(bminy, st.p.y, fma_rn(bmaxy, st.q.y, st.cn.y));
    const float tnz = fma_rn(bminz, st.p.z, fma_rn(bmaxz, st.q.z, st.cn.z));
    const float tfx = fma_rn(bminx, st.q.x, fma_rn(bmaxx, st.p.x, st.cf.x));
    const float tfy = fma_rn(bminy, st.q.y, fma_rn(bmaxy, st.p.y, st.cf.y));
    const float tfz = fma_rn(bminz, st.q.z, fma_rn(bmaxz, st.p.z, st.cf.z));
    // The entry distance is NOT clamped at 0 (a box around the origin gets a negative key and goes
    // first); boxes behind the origin are rejected by tfar >= 0 instead.  An empty slot's NaN box
    // gives NaN on every axis: fmaxf / fminf drop NaN operands, so tnear = NaN (compares false)
    // while tfar = tlimit.
    const float tnear = fmaxf(fmaxf(tnx, tny), tnz);
    const float tfar = fminf(fminf(tfx, tfy), fminf(tfz, tlimit));
    return (tnear <= tfar && tfar >= 0.0f) ? tnear : u2f(0x7F800000u);
}

// the reference's test of one child (world space: the TLAS), same return convention
SPB_HD float slab_key(float bminx, float bminy, float bminz, float bmaxx, float bmaxy, float bmaxz, f3 o, f3 inv, float tlimit)
{
    float tnear;
    const bool hit = slab_fast(bminx, bminy, bminz, bmaxx, bmaxy, bmaxz, o, inv, tnear);
    return (hit && tnear <= tlimit) ? tnear : u2f(0x7F800000u);
}

SPB_HD void trav2_push(Trav2 &st, TravEntry *stack, uint32_t ref, float tnear)
{
    TravEntry e;
    e.ref = ref;
    e.tnear = tnear;
    stack[st.sp] = e;
    st.sp++;
}

// pops until an entry passes the cull distance; the sentinels (tnear = -inf) always do
template <bool CULL>
SPB_HD void trav2_pop(Trav2 &st, const TravEntry *stack)
{
    for (;;)
    {
        st.sp--;
        TravEntry e = load_entry(stack + st.sp);
        if (CULL && !(e.tnear <= st.tcull)) continue;
        st.cur = e.ref;
        return;
    }
}

// Start of a query: world ray (o, d).  Initialises the lane's record; returns false when there is
// nothing to walk: st.cur = DONE, with T2_SLOW set if the caller must finish the ray with
// intersect_scene().  On true the record holds the world ray and its reciprocal.
template <int STRIDE>
SPB_HD bool trav2_start(const DScene &S, f3 o, f3 d, Trav2 &st, const T2View<STRIDE> &v)
{
    st.sp = 0;
    st.tcull = u2f(0x7F800000u);
    st.cur = SPB_NODE_DONE;
    st.object = SPB_T2_NO_OBJECT;
    v.f(T2_BT) = -1.0f;
    v.u(T2_BSLOT) = 0;
    v.u(T2_BOBJECT) = 0xFFFFFFFFu;
    v.u(T2_SLOW) = 0;
    if (S.tlasRoot == SPB_REF_EMPTY) return false;
    if (!trav2_ray_ok(o, d, S.tlasExtent))
    {
        v.u(T2_SLOW) = 1;
        return false;
    }
    v.set3(T2_WOX, o);
    v.set3(T2_WDX, d);
    v.set3(T2_WIX, mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z));
    return true;
}

// General scenes: the walk starts at the TLAS root.
template <int STRIDE>
SPB_HD void trav2_begin(const DScene &S, f3 o, f3 d, Trav2 &st, const T2View<STRIDE> &v, TravEntry *stack)
{
    if (!trav2_start(S, o, d, st, v)) return;
    trav2_push(st, stack, SPB_NODE_DONE, u2f(0xFF800000u));
    st.cur = S.tlasRoot;
}

// NODE step: st.cur is a node index.  `tlas`: optional copy of the first `tlasCount` TLAS nodes in
// shared memory (9 float4 per node: 8 + one of padding, so that a quarter-warp reading the same
// word of eight different nodes hits eight different bank groups); null = read from global memory.
template <bool CULL, int STRIDE>
SPB_HD void trav2_node(const DScene &S, Trav2 &st, const T2View<STRIDE> &v, TravEntry *stack, Counters *counters,
                       const v4f *tlas = nullptr, uint32_t tlasCount = 0)
{
    const float inf = u2f(0x7F800000u);
    v4f minx, miny, minz, maxx, maxy, maxz, refsf, meta;
    const uint32_t local = st.cur - S.tlasRoot;
    if (tlas && st.object == SPB_T2_NO_OBJECT && local < tlasCount)
    {
        const v4f *n = tlas + (size_t)local * 9;
        minx = n[0]; miny = n[1]; minz = n[2]; maxx = n[3]; maxy = n[4]; maxz = n[5]; refsf = n[6];
    }
    else
    {
        const v4f *n = S.nodes + (size_t)st.cur * 8;
        ld8(n + 0, minx, miny);
        ld8(n + 2, minz, maxx);
        ld8(n + 4, maxy, maxz);
        ld8(n + 6, refsf, meta);
    }
    if (counters) counters->nodeVisits++;

    const float tlimit = CULL ? st.tcull : inf;
    float k0, k1, k2, k3;
    if (st.object != SPB_T2_NO_OBJECT)
    {
        k0 = slab_wide(minx.x, miny.x, minz.x, maxx.x, maxy.x, maxz.x, st, tlimit);
        k1 = slab_wide(minx.y, miny.y, minz.y, maxx.y, maxy.y, maxz.y, st, tlimit);
        k2 = slab_wide(minx.z, miny.z, minz.z, maxx.z, maxy.z, maxz.z, st, tlimit);
        k3 = slab_wide(minx.w, miny.w, minz.w, maxx.w, maxy.w, maxz.w, st, tlimit);
    }
    else
    {
        const f3 o = v.f3at(T2_WOX), inv = v.f3at(T2_WIX);
        k0 = slab_key(minx.x, miny.x, minz.x, maxx.x, maxy.x, maxz.x, o, inv, tlimit);
        k1 = slab_key(minx.y, miny.y, minz.y, maxx.y, maxy.y, maxz.y, o, inv, tlimit);
        k2 = slab_key(minx.z, miny.z, minz.z, maxx.z, maxy.z, maxz.z, o, inv, tlimit);
        k3 = slab_key(minx.w, miny.w, minz.w, maxx.w, maxy.w, maxz.w, o, inv, tlimit);
    }
    uint32_t r0 = f2u(refsf.x), r1 = f2u(refsf.y), r2 = f2u(refsf.z), r3 = f2u(refsf.w);
    if (CULL)
    {
        sortx(k0, r0, k1, r1);
        sortx(k2, r2, k3, r3);
        sortx(k0, r0, k2, r2);
        sortx(k1, r1, k3, r3);
        sortx(k1, r1, k2, r2);
    }
    else
    {
        if (!(k0 < inf)) { k0 = k1; r0 = r1; k1 = inf; }
        if (!(k1 < inf)) { k1 = k2; r1 = r2; k2 = inf; }
        if (!(k2 < inf)) { k2 = k3; r2 = r3; k3 = inf; }
        if (!(k0 < inf)) { k0 = k1; r0 = r1; k1 = inf; }
        if (!(k1 < inf)) { k1 = k2; r1 = r2; k2 = inf; }
        if (!(k0 < inf)) { k0 = k1; r0 = r1; k1 = inf; }
    }
    if (k3 < inf) trav2_push(st, stack, r3, k3);
    if (k2 < inf) trav2_push(st, stack, r2, k2);
    if (k1 < inf) trav2_push(st, stack, r1, k1);
    if (k0 < inf)
    {
        st.cur = r0;
        return;
    }
    trav2_pop<CULL>(st, stack);
}

// Object entry (sp_scene.cpp:274-276) for object `index`, whose own world box passed the exact
// test.  The world ray is in the record; the walk continues in object space above `sentinel`.
// Returns false (T2_SLOW set, st.cur = DONE) when the object-space ray needs the exact walk.
template <bool CULL, int STRIDE>
SPB_HD bool trav2_enter(const DScene &S, uint32_t index, uint32_t meshRoot, Trav2 &st, const T2View<STRIDE> &v,
                        TravEntry *stack, uint32_t sentinel)
{
    const float inf = u2f(0x7F800000u), ninf = u2f(0xFF800000u);
    const f3 wo = v.f3at(T2_WOX), wd = v.f3at(T2_WDX);
    m4 invModel = load_m4(S.objInv + (size_t)index * 4);
    f3 lo = xform(invModel, wo, 1.0f);
    float scaleLen;
    f3 ld = normalize3(xform(invModel, wd, 0.0f), &scaleLen);
    const float extent = ld4(S.objBox + (size_t)index * 2).w;
    if (!trav2_ray_ok(lo, ld, extent))
    {
        v.u(T2_SLOW) = 1;
        st.cur = SPB_NODE_DONE;
        return false;
    }
    v.f(T2_WORLDCULL) = st.tcull;
    st.tcull = inf;
    const float bT = v.f(T2_BT);
    if (CULL && bT >= 0.0f) st.tcull = (bT + cull_pad(wo, bT)) * scaleLen * SPB_CULL_SLACK;
    const f3 inv = mk3(1.0f / ld.x, 1.0f / ld.y, 1.0f / ld.z);
    v.set3(T2_OX, lo);
    v.set3(T2_DX, ld);
    v.set3(T2_IX, inv);
    v.f(T2_LT) = -1.0f;
    v.u(T2_LSLOT) = 0;
    st.object = index;
    trav2_constants(lo, inv, extent, st);
    trav2_push(st, stack, sentinel, ninf);
    st.cur = meshRoot;
    return true;
}

// The object's closest hit carried back to world space (sp_scene.cpp:296-322); updates the closest
// hit overall and returns the world cull distance that follows from it.
template <int STRIDE>
SPB_HD float trav2_leave(const DScene &S, uint32_t object, float worldCull, const T2View<STRIDE> &v)
{
    const float lT = v.f(T2_LT);
    if (lT >= 0.0f)
    {
        const f3 wo = v.f3at(T2_WOX), wd = v.f3at(T2_WDX);
        m4 model = load_m4(S.objModel + (size_t)object * 4);
        f3 localHit = add3(v.f3at(T2_OX), mul3(v.f3at(T2_DX), lT));
        f3 worldHit = xform(model, localHit, 1.0f);
        float t = dot3(sub3(worldHit, wo), wd);
        float bT = v.f(T2_BT);
        if (t < bT || bT < 0.0f)
        {
            v.f(T2_BT) = t;
            v.u(T2_BOBJECT) = object;
            v.u(T2_BSLOT) = v.u(T2_LSLOT);
            float c2 = (t + cull_pad(wo, t)) * SPB_CULL_SLACK;
            if (t > 0.0f && c2 < worldCull) worldCull = c2;
        }
    }
    return worldCull;
}

// EXIT step (st.cur == SPB_NODE_EXIT): leave the object and go on with the TLAS entries below.
template <bool CULL, int STRIDE>
SPB_HD void trav2_exit(const DScene &S, Trav2 &st, const T2View<STRIDE> &v, const TravEntry *stack)
{
    st.tcull = trav2_leave(S, st.object, v.f(T2_WORLDCULL), v);
    st.object = SPB_T2_NO_OBJECT;
    trav2_pop<CULL>(st, stack);
}

// LEAF step: st.cur is SPB_REF_LEAF | slot -- inside an object a triangle whose box passed the
// conservative test of its parent's node step; in the TLAS an object whose own box passed the
// exact one.
template <bool CULL, int STRIDE>
SPB_HD void trav2_leaf(const DScene &S, Trav2 &st, const T2View<STRIDE> &v, TravEntry *stack, Counters *counters)
{
    const uint32_t index = st.cur & ~SPB_REF_LEAF;
    if (st.object != SPB_T2_NO_OBJECT)
    {
        const f3 o = v.f3at(T2_OX), inv = v.f3at(T2_IX);
        const v4f *tp = S.tris + (size_t)index * 3;
        v4f a = ld4(tp + 0), b = ld4(tp + 1), c = ld4(tp + 2);
        // the triangle's own box (sp_scene.cpp:35-50) and the reference's test of it (bvh.cpp:236-255)
        float tn;
        const bool own = slab_fast(fminf(a.x, fminf(b.x, c.x)), fminf(a.y, fminf(b.y, c.y)), fminf(a.z, fminf(b.z, c.z)),
                                   fmaxf(a.x, fmaxf(b.x, c.x)), fmaxf(a.y, fmaxf(b.y, c.y)), fmaxf(a.z, fmaxf(b.z, c.z)),
                                   o, inv, tn);
        if (own && (!CULL || tn <= st.tcull))
        {
            if (counters) counters->triangleTests++;
            const f3 d = v.f3at(T2_DX);
            float t, uu, vv;
            if (ray_triangle_mt(o, d, mk3(a.x, a.y, a.z), mk3(b.x, b.y, b.z), mk3(c.x, c.y, c.z), t, uu, vv))
            {
                const float lT = v.f(T2_LT);
                if (t > 0.0f && (t < lT || lT < 0.0f))
                {
                    v.f(T2_LT) = t;
                    v.u(T2_LSLOT) = index;
                    float c2 = t * SPB_CULL_SLACK;
                    if (c2 < st.tcull) st.tcull = c2;
                }
            }
        }
    }
    else
    {
        v4u info = ld4u(S.objInfo + index);
        if (counters) counters->objectTests++;
        if (info.x != SPB_REF_EMPTY)
        {
            trav2_enter<CULL>(S, index, info.x, st, v, stack, SPB_NODE_EXIT);
            return;
        }
    }
    trav2_pop<CULL>(st, stack);
}

// Single-object scenes (C1-C4): what the walk would do before it reaches the mesh tree -- the TLAS
// root's one child is the object: its own box, exactly; then the object entry -- done at once when
// the ray starts (every lane of a refilling warp is busy here, and two rounds of the step loop are
// saved).  Leaves st.cur = the mesh root, or DONE.  The sentinel under the object's entries is DONE
// rather than EXIT: the walk ends inside the object, and trav2_finish() carries the hit out.
template <bool CULL, int STRIDE>
SPB_HD void trav2_begin_single(const DScene &S, f3 o, f3 d, Trav2 &st, const T2View<STRIDE> &v, TravEntry *stack,
                               Counters *counters)
{
    if (!trav2_start(S, o, d, st, v)) return;
    if (counters) counters->nodeVisits++;
    v4f bmin = ld4(S.objBox), bmax = ld4(S.objBox + 1);
    float tn;
    if (!slab_fast(bmin.x, bmin.y, bmin.z, bmax.x, bmax.y, bmax.z, o, v.f3at(T2_WIX), tn)) return;
    v4u info = ld4u(S.objInfo);
    if (counters) counters->objectTests++;
    if (info.x == SPB_REF_EMPTY) return;
    trav2_enter<CULL>(S, 0, info.x, st, v, stack, SPB_NODE_DONE);
}

// End of a query: the closest hit.  (A single-object walk ends inside the object.)
template <int STRIDE>
SPB_HD Hit trav2_finish(const DScene &S, const Trav2 &st, const T2View<STRIDE> &v)
{
    if (st.object != SPB_T2_NO_OBJECT && !v.u(T2_SLOW)) trav2_leave(S, st.object, 0.0f, v);
    Hit h;
    h.t = v.f(T2_BT);
    h.object = (int32_t)v.u(T2_BOBJECT);
    h.slot = v.u(T2_BSLOT);
    h.u = h.v = 0.0f;
    h.localOrigin = h.localDirection = mk3(0, 0, 0);
    h.localT = -1.0f;
    return h;
}

// The second machine run to completion for one ray (what each lane of k_trace does); hostsim
// checks it against intersect_scene() and the oracle.
template <bool CULL>
SPB_HD Hit intersect_scene_stepped2(const DScene &S, f3 o, f3 d, uint32_t *stack, float *stackT, Counters *counters)
{
    Trav2 st;
    float record[T2_WORDS];
    T2View<1> v;
    v.base = record;
    TravEntry entries[SPB_STACK_SIZE];
    const bool single = S.objectCount == 1;
    if (single) trav2_begin_single<CULL>(S, o, d, st, v, entries, counters);
    else trav2_begin(S, o, d, st, v, entries);
    while (st.cur != SPB_NODE_DONE)
    {
        if (st.cur == SPB_NODE_EXIT) trav2_exit<CULL>(S, st, v, entries);
        else if ((st.cur & SPB_REF_LEAF) == 0) trav2_node<CULL>(S, st, v, entries, counters);
        else trav2_leaf<CULL>(S, st, v, entries, counters);
    }
    if (v.u(T2_SLOW)) return intersect_scene<CULL, false>(S, o, d, stack, stackT, counters);
    Hit h = trav2_finish(S, st, v);
    if (h.object >= 0) hit_barycentrics(S, o, d, h);
    return h;
}

struct Surface
{
    f3 normal;       // world normal (sp_scene.cpp:309-312)
    float uvx, uvy;  // interpolated texture coordinates (sp_scene.cpp:199-205)
    uint32_t material;
    uint32_t triangle; // triangle index within its mesh
};

// Attributes of the winning triangle (sp_scene.cpp:196-216, 309-312); recomputing e1 x e2 from
// the same vertices gives the same bits as the value Moller-Trumbore returned.
SPB_HD Surface resolve_hit(const DScene &S, const Hit &hit)
{
    Surface s;
    v4u info = ld4u(S.objInfo + hit.object);
    const v4f *tp = S.tris + (size_t)hit.slot * 3;
    v4f a = ld4(tp + 0), b = ld4(tp + 1), c = ld4(tp + 2);
    uint32_t tri = f2u(a.w);
    const v4f *sp = S.shade + ((size_t)info.y + tri) * 4;
    v4f s0 = ld4(sp + 0), s1 = ld4(sp + 1), s2 = ld4(sp + 2), s3 = ld4(sp + 3);

    float u = hit.u, v = hit.v;
    float w = 1.0f - u - v;
    // uv = uv0 * w + uv1 * u + uv2 * v
    s.uvx = s0.w * w + s2.w * u + s3.y * v;
    s.uvy = s1.w * w + s3.x * u + s3.z * v;

    f3 localNormal;
    if (info.z != 0)
    {
        f3 n0 = mk3(s0.x, s0.y, s0.z), n1 = mk3(s1.x, s1.y, s1.z), n2 = mk3(s2.x, s2.y, s2.z);
        localNormal = normalize3(add3(add3(mul3(n0, w), mul3(n1, u)), mul3(n2, v)));
    }
    else
    {
        f3 pa = mk3(a.x, a.y, a.z);
        localNormal = cross3(sub3(mk3(b.x, b.y, b.z), pa), sub3(mk3(c.x, c.y, c.z), pa));
    }
    m4 model = load_m4(S.objModel + (size_t)hit.object * 4);
    s.normal = normalize3(xform(model, localNormal, 0.0f));
    s.material = info.w;
    s.triangle = tri;
    return s;
}

// ------------------------------------------------------------------------------------------
// materials and shading

// SampleImageNearest (image.h:3-18).  The reference does not clamp; an index past the end of
// the image (v == 1 exactly) is undefined there, clamped to the last texel and counted here.
SPB_HD f3 sample_nearest(const DImage &img, float u, float v, Counters *counters)
{
    float fx = u * (float)img.width;
    float fy = v * (float)img.height;
    float flx = floorf(fx), fly = floorf(fy);
    // (u32) of a negative float is undefined in C; saturate at 0
    uint32_t x = flx > 0.0f ? (flx < 4294967040.0f ? (uint32_t)flx : 0xFFFFFF00u) : 0u;
    uint32_t y = fly > 0.0f ? (fly < 4294967040.0f ? (uint32_t)fly : 0xFFFFFF00u) : 0u;
    uint64_t index = (uint64_t)y * img.width + x;
    uint64_t last = (uint64_t)img.width * img.height - 1;
    if (index > last)
    {
        index = last;
        if (counters) counters->envClamped++;
    }
    v4f p = ld4(img.pixels + index);
    return mk3(p.x, p.y, p.z);
}

// SampleImageBilinear (image.h:34-73), clamp-to-edge, Lerp(a,b,t) = a*(1-t) + b*t
SPB_HD f3 sample_bilinear(const DImage &img, float u, float v)
{
    float px = (u * (float)img.width) - 0.5f;
    float py = (v * (float)img.height) - 0.5f;
    px = rmax(px, 0.0f);
    py = rmax(py, 0.0f);
    uint32_t x0 = (uint32_t)floorf(px), y0 = (uint32_t)floorf(py);
    if (x0 > img.width - 1) x0 = img.width - 1;
    if (y0 > img.height - 1) y0 = img.height - 1;
    uint32_t x1 = x0 + 1, y1 = y0 + 1;
    float fx = px - (float)x0, fy = py - (float)y0;
    if (x1 > img.width - 1) x1 = img.width - 1;
    if (y1 > img.height - 1) y1 = img.height - 1;
    v4f s0 = ld4(img.pixels + (size_t)y0 * img.width + x0);
    v4f s1 = ld4(img.pixels + (size_t)y0 * img.width + x1);
    v4f s2 = ld4(img.pixels + (size_t)y1 * img.width + x0);
    v4f s3 = ld4(img.pixels + (size_t)y1 * img.width + x1);
    float ax = 1.0f - fx, ay = 1.0f - fy;
    f3 t0 = mk3(s0.x * ax + s1.x * fx, s0.y * ax + s1.y * fx, s0.z * ax + s1.z * fx);
    f3 t1 = mk3(s2.x * ax + s3.x * fx, s2.y * ax + s3.y * fx, s2.z * ax + s3.z * fx);
    return mk3(t0.x * ay + t1.x * fy, t0.y * ay + t1.y * fy, t0.z * ay + t1.z * fy);
}

struct MaterialOut { f3 albedo, emission; float roughness; };

// ToSphericalCoordinates (math_lib.h:849-860), MapToEquirectangular (math_lib.h:873-884), then
// uv.y = 1 - uv.y (sp_material_system.cpp:88-93)
template <int MATH>
SPB_HD void equirect_uv(f3 dir, float &eu, float &ev)
{
    float inc = m_atan2<MATH>(sqrtf(dir.x * dir.x + dir.z * dir.z), dir.y);
    float az = m_atan2<MATH>(dir.z, dir.x);
    if (az < 0.0f) az += 2.0f * SPB_PI;
    eu = az / (2.0f * SPB_PI);
    ev = m_cos<MATH>(inc) * 0.5f + 0.5f;
    ev = 1.0f - ev;
}

// sp_FindMaterialById + sp_EvaluateMaterial (sp_material_system.cpp:15-29, 59-105) and the
// missing-material fallback of ComputeRadianceForPath (simd_path_tracer.cpp:122-133)
template <int MATH, int ENVFILTER>
SPB_HD MaterialOut evaluate_material(const DMaterials &M, uint32_t materialId, f3 outgoingDir,
                                     float uvx, float uvy, Counters *counters)
{
    MaterialOut out;
    out.albedo = mk3(0, 0, 0);
    out.emission = mk3(0, 0, 0);
    out.roughness = 0.0f;
    int slot = -1;
    for (uint32_t i = 0; i < M.count; ++i)
    {
        if (M.keys[i] == materialId)
        {
            slot = (int)i;
            break;
        }
    }
    if (slot < 0)
    {
        out.emission = mk3(1.0f, 0.0f, 1.0f);
        return out;
    }
    int ai = M.albedoImage[slot];
    if (ai >= 0)
        out.albedo = sample_nearest(M.images[ai], uvx, uvy, counters);
    else
        out.albedo = mk3(M.albedo[slot][0], M.albedo[slot][1], M.albedo[slot][2]);

    int ei = M.emissionImage[slot];
    if (ei >= 0)
    {
        // ToSphericalCoordinates(-outgoingDir) (math_lib.h:849-860), MapToEquirectangular
        // (math_lib.h:873-884), then uv.y = 1 - uv.y (sp_material_system.cpp:88-93)
        float eu, ev;
        equirect_uv<MATH>(neg3(outgoingDir), eu, ev);
        if (ENVFILTER == 0)
            out.emission = sample_nearest(M.images[ei], eu, ev, counters);
        else
            out.emission = sample_bilinear(M.images[ei], eu, ev);
    }
    else
        out.emission = mk3(M.emission[slot][0], M.emission[slot][1], M.emission[slot][2]);
    out.roughness = M.roughness[slot];
    return out;
}

// One iteration of ComputeRadianceForPath (simd_path_tracer.cpp:113-171) split in two: the part
// that does not depend on the radiance arriving from the next vertex (emission E, BRDF weight
// W = kD*albedo/pi + specular, cosine) is evaluated when the vertex is created ...
struct VertexTerms { f3 E, W; float cosine; };

template <int MATH, int ENVFILTER>
SPB_HD VertexTerms vertex_terms(const DMaterials &M, uint32_t materialId, f3 L, f3 N, f3 V,
                                float uvx, float uvy, Counters *counters)
{
    MaterialOut mo = evaluate_material<MATH, ENVFILTER>(M, materialId, V, uvx, uvy, counters);
    VertexTerms vt;
    vt.E = mo.emission;
    vt.cosine = rmax(0.0f, dot3(N, L));

    f3 H = normalize3(add3(L, V));
    // FresnelSchlick(Max(Dot(H,V),0), F0 = 0.04) (simd_path_tracer.cpp:65-70)
    float p5 = m_pow5<MATH>(1.0f - rmax(dot3(H, V), 0.0f));
    float Fx = 0.04f + (1.0f - 0.04f) * p5;
    f3 F = mk3(Fx, Fx, Fx);
    f3 kD = mk3(1.0f - F.x, 1.0f - F.y, 1.0f - F.z);
    float oneOverPI = 1.0f / SPB_PI;

    float roughness = mo.roughness;
    // DistributionGGX (:72-84)
    float a = roughness * roughness;
    float a2 = a * a;
    float NdotH = rmax(dot3(N, H), 0.0f);
    float NdotH2 = NdotH * NdotH;
    float dd = (NdotH2 * (a2 - 1.0f) + 1.0f);
    dd = SPB_PI * dd * dd;
    float NDF = a2 / dd;
    // GeometrySmith (:86-105)
    float NdotV = rmax(dot3(N, V), 0.0f);
    float NdotL = rmax(dot3(N, L), 0.0f);
    float r1 = (roughness + 1.0f);
    float k = (r1 * r1) / 8.0f;
    float ggx2 = NdotV / (NdotV * (1.0f - k) + k);
    float ggx1 = NdotL / (NdotL * (1.0f - k) + k);
    float G = ggx1 * ggx2;

    f3 numerator = mul3(F, NDF * G);
    float denominator = 4.0f * rmax(dot3(N, V), 0.0f) * rmax(dot3(N, L), 0.0f) + 0.0001f;
    f3 specular = mul3(numerator, 1.0f / denominator);
    vt.W = add3(mul3(had3(kD, mo.albedo), oneOverPI), specular);
    return vt;
}

// ... and the part that does (clamp, multiply, add emission) runs back to front afterwards.
SPB_HD f3 fold_radiance(const VertexTerms &vt, f3 incoming, float clampValue)
{
    if (clampValue > 0.0f)
    {
        incoming.x = rmin(rmax(incoming.x, 0.0f), clampValue);
        incoming.y = rmin(rmax(incoming.y, 0.0f), clampValue);
        incoming.z = rmin(rmax(incoming.z, 0.0f), clampValue);
    }
    return add3(vt.E, mul3(had3(vt.W, incoming), vt.cosine));
}

// Radiance of the vertex where a ray escapes (simd_path_tracer.cpp:290-300 creates it with
// incomingDir = normal = 0; ComputeRadianceForPath folds it first, with zero incoming radiance):
//   L = E + (W (.) clamp(0)) * cosine,  cosine = Max(0, Dot(0, 0)) = +0.
// When the background material has a constant, finite, non-negative albedo and a finite
// non-negative roughness, W is finite and non-negative for every direction (F in [0.04, 1],
// D = a2 / pi, G = 0 / k with k = (r + 1)^2 / 8 > 0), so W * 0 * 0 = +0 and L = E + (+0) in every
// component: exactly E, with a -0 turned into +0 -- which is what `E + 0.0f` computes, NaN and inf
// included.  The Fresnel / GGX / Smith terms (a pow and a dozen divisions per escaped ray) are then
// dead arithmetic.  Any other background material takes the full evaluation.
template <int MATH, int ENVFILTER>
SPB_HD f3 miss_radiance(const DMaterials &M, f3 V, float clampValue, Counters *counters)
{
    f3 zero = mk3(0.0f, 0.0f, 0.0f);
    if (M.simpleBackground)
    {
        MaterialOut mo = evaluate_material<MATH, ENVFILTER>(M, M.backgroundId, V, 0.0f, 0.0f, counters);
        return mk3(mo.emission.x + 0.0f, mo.emission.y + 0.0f, mo.emission.z + 0.0f);
    }
    VertexTerms vt = vertex_terms<MATH, ENVFILTER>(M, M.backgroundId, zero, zero, V, 0.0f, 0.0f, counters);
    return fold_radiance(vt, zero, clampValue);
}

// RandomDirectionOnHemisphere (math_lib.h:932-946)
template <int MATH>
SPB_HD f3 random_hemisphere(f3 normal, uint32_t &rng)
{
    float theta = SPB_PI * rand_unilateral(rng);
    float phi = SPB_PI * rand_bilateral(rng);
    float st = m_sin<MATH>(theta);
    f3 dir;
    dir.x = st * m_cos<MATH>(phi);
    dir.z = st * m_sin<MATH>(phi);
    dir.y = m_cos<MATH>(theta);
    if (dot3(dir, normal) < 0.0f) dir = neg3(dir);
    return dir;
}

// ------------------------------------------------------------------------------------------
// Shortcuts of the wavefront renderer that rest on the size of the reference's jitter
// (+-0.5/width of a PIXEL, simd_path_tracer.cpp:222-226): written here, for host and device, so
// that tests/hostsim can check them against the plain evaluation without a GPU.

// Coverage of one instanced triangle (index i over all objects' triangles, DScene::objTris):
// object-space vertices -> world (model matrix) -> film (sp_CalculateFilmPositions inverted,
// simd_path_tracer.cpp:38-63) in double; the pixel bounding box of the three projections, padded
// by 2 pixels, marks the 8x4 pixel blocks of the strip [x0,x1) x [y0,y1) it touches in
// coverage[blocksX * blocksY]; coverage[blocksX * blocksY] is the "everything" flag, raised when
// the triangle reaches the camera plane (its projection is unbounded).
SPB_HD void cover_triangle(const DScene &S, const DCamera &c, unsigned long long i, uint32_t x0, uint32_t y0,
                           uint32_t x1, uint32_t y1, uint32_t blocksX, uint32_t blocksY, uint8_t *coverage)
{
    const double Fx = (double)c.filmCenter.x - c.position.x, Fy = (double)c.filmCenter.y - c.position.y,
                 Fz = (double)c.filmCenter.z - c.position.z;
    const double dist = sqrt(Fx * Fx + Fy * Fy + Fz * Fz);
    const uint32_t *first = S.objTris, *prefix = S.objTris + S.objectCount;
    const uint32_t blocks = blocksX * blocksY;
    // object of this instanced triangle: last prefix entry <= i
    uint32_t lo = 0, hi = S.objectCount;
    while (hi - lo > 1)
    {
        uint32_t mid = (lo + hi) >> 1;
        if (prefix[mid] <= i) lo = mid; else hi = mid;
    }
    const uint32_t object = lo;
    const size_t slot = (size_t)first[object] + (size_t)(i - prefix[object]);
    const v4f *tp = S.tris + slot * 3;
    const v4f *mp = S.objModel + (size_t)object * 4;
    const v4f m0 = mp[0], m1 = mp[1], m2 = mp[2], m3 = mp[3];
    double px[3], py[3], depthMin = 1e300, depthMax = -1e300, scale = 0.0;
    for (int k = 0; k < 3; ++k)
    {
        v4f v = tp[k];
        double wx = (double)m0.x * v.x + (double)m1.x * v.y + (double)m2.x * v.z + m3.x - c.position.x;
        double wy = (double)m0.y * v.x + (double)m1.y * v.y + (double)m2.y * v.z + m3.y - c.position.y;
        double wz = (double)m0.z * v.x + (double)m1.z * v.y + (double)m2.z * v.z + m3.z - c.position.z;
        double depth = (wx * Fx + wy * Fy + wz * Fz) / dist;
        double lambda = depth / dist; // w = lambda * (film point - camera position)
        double fa = (wx * c.right.x + wy * c.right.y + wz * c.right.z) / lambda / c.halfFilmWidth;
        double fb = (wx * c.up.x + wy * c.up.y + wz * c.up.z) / lambda / c.halfFilmHeight;
        px[k] = (fa + 1.0) * 0.5 * c.width;
        py[k] = (1.0 - (fb + 1.0) * 0.5) * c.height;
        depthMin = fmin(depthMin, depth);
        depthMax = fmax(depthMax, depth);
        scale = fmax(scale, fabs(wx) + fabs(wy) + fabs(wz));
    }
    if (!(depthMax > 0.0)) return; // entirely behind the camera plane: no camera ray goes there
    if (!(depthMin > 1e-6 * scale) || !(dist > 0.0))
    {
        coverage[blocks] = 1; // reaches the camera plane: its projection is unbounded
        return;
    }
    double xlo = fmin(px[0], fmin(px[1], px[2])) - 2.0, xhi = fmax(px[0], fmax(px[1], px[2])) + 2.0;
    double ylo = fmin(py[0], fmin(py[1], py[2])) - 2.0, yhi = fmax(py[0], fmax(py[1], py[2])) + 2.0;
    if (!(xlo == xlo) || !(xhi == xhi) || !(ylo == ylo) || !(yhi == yhi)) { coverage[blocks] = 1; return; }
    if (xhi < (double)x0 || yhi < (double)y0 || xlo >= (double)x1 || ylo >= (double)y1) return;
    uint32_t bx0 = xlo <= (double)x0 ? 0u : ((uint32_t)xlo - x0) >> 3;
    uint32_t by0 = ylo <= (double)y0 ? 0u : ((uint32_t)ylo - y0) >> 2;
    uint32_t bx1 = xhi >= (double)(x1 - 1) ? blocksX - 1 : ((uint32_t)xhi - x0) >> 3;
    uint32_t by1 = yhi >= (double)(y1 - 1) ? blocksY - 1 : ((uint32_t)yhi - y0) >> 2;
    for (uint32_t by = by0; by <= by1; ++by)
        for (uint32_t bx = bx0; bx <= bx1; ++bx) coverage[by * blocksX + bx] = 1;
}

// jitter-free camera ray through the centre of pixel (x, y): sp_CalculateFilmPositions
// (simd_path_tracer.cpp:38-63) on (x + 0.5, y + 0.5)
SPB_HD f3 centre_direction(const DCamera &cam, uint32_t x, uint32_t y)
{
    float fx = ((float)x + 0.5f) / (float)cam.width, fy = 1.0f - ((float)y + 0.5f) / (float)cam.height;
    fx = fx * 2.0f - 1.0f;
    fy = fy * 2.0f - 1.0f;
    f3 filmP = add3(add3(mul3(cam.right, cam.halfFilmWidth * fx), mul3(cam.up, cam.halfFilmHeight * fy)), cam.filmCenter);
    return normalize3(sub3(filmP, cam.position));
}

// Min(v0, Min(v1, v2)) / Max(v0, Max(v1, v2)) with the ternaries of build_mesh_accel(): the box
// of a leaf, bit for bit
SPB_HD void triangle_box(const v4f &a, const v4f &b, const v4f &c, f3 &lo, f3 &hi)
{
    float l, h;
    l = b.x < c.x ? b.x : c.x; lo.x = a.x < l ? a.x : l;
    l = b.y < c.y ? b.y : c.y; lo.y = a.y < l ? a.y : l;
    l = b.z < c.z ? b.z : c.z; lo.z = a.z < l ? a.z : l;
    h = b.x > c.x ? b.x : c.x; hi.x = a.x > h ? a.x : h;
    h = b.y > c.y ? b.y : c.y; hi.y = a.y > h ? a.y : h;
    h = b.z > c.z ? b.z : c.z; hi.z = a.z > h ? a.z : h;
}

// triangle slots kept per pixel; a pixel whose padded centre ray enters more leaf boxes falls back
#define SPB_CAND_MAX 23u
#define SPB_CAND_STRIDE (SPB_CAND_MAX + 1u)
#define SPB_CAND_FALLBACK 0xFFFFFFFFu

// Candidate triangles of a pixel (single-object scenes).  The 64 camera rays of a pixel differ by
// about 1e-7 rad; walking the tree once per SAMPLE repeats the same box tests 64 times.  This
// walks it once per PIXEL with the jitter-free centre ray against boxes padded by
// delta = 1e-4 x (mesh extent + |origin|) -- a thousand times the separation of the pixel's rays
// inside the scene plus every rounding of the transforms and slab products -- and lists every
// triangle whose padded box the centre ray enters, without any distance culling.  Whatever sample
// ray passes the exact slab test of a triangle's own box (the reference's per-leaf test,
// bvh.cpp:236-255) therefore finds that triangle in the list.  out[0] = count or
// SPB_CAND_FALLBACK (list overflow, non-finite reciprocal direction), out[1..] = triangle slots.
SPB_HD void collect_candidates(const DScene &S, const DCamera &cam, uint32_t x, uint32_t y, uint32_t *out)
{
    f3 o = cam.position, d = centre_direction(cam, x, y);
    v4u info = ld4u(S.objInfo);
    m4 invModel = load_m4(S.objInv);
    f3 lo = xform(invModel, o, 1.0f);
    f3 ld = normalize3(xform(invModel, d, 0.0f));
    if (info.x == SPB_REF_EMPTY) { out[0] = 0; return; }
    if (any_nonfinite_inv(d) || any_nonfinite_inv(ld)) { out[0] = SPB_CAND_FALLBACK; return; }
    f3 inv = mk3(1.0f / ld.x, 1.0f / ld.y, 1.0f / ld.z);
    // padding from the mesh's extent (its root node's child boxes) and the origin
    float delta;
    {
        const v4f *n = S.nodes + (size_t)info.x * 8;
        float big = 0.0f;
        for (int k = 0; k < 6; ++k)
        {
            v4f v = ld4(n + k);
            // fmaxf skips the NaN of empty slots
            big = fmaxf(big, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
        }
        delta = 1.0e-4f * (6.0f * big + fabsf(lo.x) + fabsf(lo.y) + fabsf(lo.z));
    }
    uint32_t stack[64];
    int sp = 0;
    uint32_t node = info.x, count = 0;
    bool overflow = false;
    for (;;)
    {
        const v4f *n = S.nodes + (size_t)node * 8;
        v4f mnx = ld4(n + 0), mny = ld4(n + 1), mnz = ld4(n + 2), mxx = ld4(n + 3), mxy = ld4(n + 4), mxz = ld4(n + 5);
        v4u refs = ld4u((const v4u *)(n + 6));
        const float bmn[4][3] = {{mnx.x, mny.x, mnz.x}, {mnx.y, mny.y, mnz.y}, {mnx.z, mny.z, mnz.z}, {mnx.w, mny.w, mnz.w}};
        const float bmx[4][3] = {{mxx.x, mxy.x, mxz.x}, {mxx.y, mxy.y, mxz.y}, {mxx.z, mxy.z, mxz.z}, {mxx.w, mxy.w, mxz.w}};
        const uint32_t ref[4] = {refs.x, refs.y, refs.z, refs.w};
        for (int k = 0; k < 4; ++k)
        {
            float tn;
            if (ref[k] == SPB_REF_EMPTY) continue;
            if (!slab_fast(bmn[k][0] - delta, bmn[k][1] - delta, bmn[k][2] - delta, bmx[k][0] + delta,
                           bmx[k][1] + delta, bmx[k][2] + delta, lo, inv, tn))
                continue;
            if (ref[k] & SPB_REF_LEAF)
            {
                if (count < SPB_CAND_MAX) out[1 + count] = ref[k] & ~SPB_REF_LEAF;
                else overflow = true;
                count++;
            }
            else if (sp < 64) stack[sp++] = ref[k];
            else overflow = true;
        }
        if (sp == 0 || overflow) break;
        node = stack[--sp];
    }
    out[0] = overflow ? SPB_CAND_FALLBACK : count;
}

// A camera ray (world o, d; st / cold fresh from trav_begin) resolved from its pixel's candidate
// list: exactly what its walk would have evaluated -- the slab test of the object's world box,
// the ray in object space, for each listed triangle the slab test of its own box (recomputed from
// the vertices) and Moller-Trumbore, the closest hit carried back to world space -- with the
// walk's functions, so the same bits; only the winner among exactly equal t may differ, as with
// any visit order.  Leaves st.cur = DONE with the result in `cold`, or everything untouched when
// the pixel falls back to the walk.
// Returns 0: resolved (bT >= 0: closest hit, world distance bT in triangle slot bSlot of object 0;
// bT < 0: no hit), 1: the pixel falls back to the walk (nothing counted), 2: the object-space ray
// needs the exact walk.  `o, d, inv`: world ray and its (finite) reciprocal.
SPB_HD int resolve_candidates(const DScene &S, const uint32_t *list, f3 o, f3 d, f3 winv, Counters *counters, float &bT,
                              uint32_t &bSlot)
{
    bT = -1.0f;
    bSlot = 0;
    const uint32_t count = list[0];
    if (count == SPB_CAND_FALLBACK) return 1;
    // the TLAS root's test of the object's world box (what the walk would run; the other three
    // slots of the root are empty in a single-object scene)
    {
        const v4f *n = S.nodes + (size_t)S.tlasRoot * 8;
        v4f mnx = ld4(n + 0), mny = ld4(n + 1), mnz = ld4(n + 2), mxx = ld4(n + 3), mxy = ld4(n + 4), mxz = ld4(n + 5);
        float tn;
        if (counters) counters->nodeVisits++;
        if (!slab_fast(mnx.x, mny.x, mnz.x, mxx.x, mxy.x, mxz.x, o, winv, tn)) return 0; // miss
    }
    // object entry (sp_scene.cpp:274-276)
    v4u info = ld4u(S.objInfo);
    if (counters) counters->objectTests++;
    if (info.x == SPB_REF_EMPTY) return 0;
    m4 invModel = load_m4(S.objInv);
    f3 lo = xform(invModel, o, 1.0f);
    f3 ld = normalize3(xform(invModel, d, 0.0f));
    if (any_nonfinite_inv(ld)) return 2;
    f3 inv = mk3(1.0f / ld.x, 1.0f / ld.y, 1.0f / ld.z);
    float lT = -1.0f;
    uint32_t lSlot = 0;
    for (uint32_t k = 0; k < count; ++k)
    {
        const uint32_t tri = list[1 + k];
        const v4f *tp = S.tris + (size_t)tri * 3;
        v4f va = ld4(tp + 0), vb = ld4(tp + 1), vc = ld4(tp + 2);
        f3 bmn, bmx;
        triangle_box(va, vb, vc, bmn, bmx);
        float tn;
        if (!slab_fast(bmn.x, bmn.y, bmn.z, bmx.x, bmx.y, bmx.z, lo, inv, tn)) continue; // the leaf's own box
        if (counters) counters->triangleTests++;
        float t, u, v;
        if (ray_triangle_mt(lo, ld, mk3(va.x, va.y, va.z), mk3(vb.x, vb.y, vb.z), mk3(vc.x, vc.y, vc.z), t, u, v))
            if (t > 0.0f && (t < lT || lT < 0.0f))
            {
                lT = t;
                lSlot = tri;
            }
    }
    // leaving the object (sp_scene.cpp:296-322)
    if (lT >= 0.0f)
    {
        m4 model = load_m4(S.objModel);
        f3 localHit = add3(lo, mul3(ld, lT));
        f3 worldHit = xform(model, localHit, 1.0f);
        bT = dot3(sub3(worldHit, o), d);
        bSlot = lSlot;
    }
    return 0;
}

// (first machine: kept for A/B builds, -DSPB_TRAV_OLD)
SPB_HD void resolve_from_candidates(const DScene &S, const uint32_t *list, f3 o, f3 d, Trav &st, TravCold &cold,
                                    Counters *counters)
{
    if (cold.slow || st.cur == SPB_NODE_DONE) return; // non-finite reciprocal direction / empty scene
    float bT;
    uint32_t bSlot;
    const int status = resolve_candidates(S, list, o, d, st.inv, counters, bT, bSlot);
    if (status == 1) return;
    st.cur = SPB_NODE_DONE;
    if (status == 2) { cold.slow = 1; return; }
    if (bT >= 0.0f)
    {
        cold.bT = bT;
        cold.bObject = 0;
        cold.bSlot = bSlot;
    }
}

// second machine: the lane's record was initialised by trav2_start() (which returned true);
// returns false when the pixel falls back to the walk
template <int STRIDE>
SPB_HD bool resolve_from_candidates2(const DScene &S, const uint32_t *list, f3 o, f3 d, Trav2 &st, const T2View<STRIDE> &v,
                                     Counters *counters)
{
    float bT;
    uint32_t bSlot;
    const int status = resolve_candidates(S, list, o, d, v.f3at(T2_WIX), counters, bT, bSlot);
    if (status == 1) return false;
    st.cur = SPB_NODE_DONE;
    if (status == 2) { v.u(T2_SLOW) = 1; return true; }
    if (bT >= 0.0f)
    {
        v.f(T2_BT) = bT;
        v.u(T2_BOBJECT) = 0;
        v.u(T2_BSLOT) = bSlot;
    }
    return true;
}

// One-lookup sky pixels.  With a simple background (miss_radiance) a sky sample's radiance is the
// texel its direction selects, E + 0.0f.  The samples of a pixel differ from the jitter-free centre
// direction c by at most `spread` (jitter x pixel angle plus the rounding of ray generation, with
// a factor 4 -- computed by the host from the camera).  Along the reference's chain (equirect_uv,
// nearest sampling: image.h:3-18) that moves the image coordinates by at most
//   |d fx| <= W * ((1.5 spread / r + 4e-7) / (2 pi) + 1.8e-7),  r = sqrt(cx^2 + cz^2)  (atan2(z, x), +2 pi, / 2 pi, * W)
//   |d fy| <= H * (spread + 5.2e-7)                                                    (atan2(r, y), cos, * 0.5 + 0.5, 1 -, * H)
// roundings of the single-precision steps included (the transcendental steps are correctly rounded
// in deterministic-math mode, the only mode this path is used in).  If the centre's coordinates are
// further than TWICE those bounds from the next texel boundary, every sample reads the centre's
// texel, and the pixel is spp times the same addition: total += E * (1 / spp), no per-sample ray at
// all.  The az = 0 seam, the poles (r < 1e-3) and the image edges sit on boundaries or are excluded
// explicitly.  Returns false when the pixel needs the sample loop.
template <int MATH, int ENVFILTER>
SPB_HD bool sky_one_lookup(const DMaterials &M, const DCamera &cam, uint32_t x, uint32_t y, float spread, uint32_t spp,
                           f3 &total)
{
    if (!(spread > 0.0f) || MATH != 0 || ENVFILTER != 0 || !M.simpleBackground) return false;
    int slot = -1;
    for (uint32_t i = 0; i < M.count; ++i)
        if (M.keys[i] == M.backgroundId) { slot = (int)i; break; }
    f3 E;
    if (slot < 0) E = mk3(1.0f, 0.0f, 1.0f);
    else if (M.emissionImage[slot] < 0) E = mk3(M.emission[slot][0], M.emission[slot][1], M.emission[slot][2]);
    else
    {
        f3 c = centre_direction(cam, x, y);
        const DImage &img = M.images[M.emissionImage[slot]];
        float eu, ev;
        equirect_uv<MATH>(c, eu, ev);
        const float W = (float)img.width, H = (float)img.height;
        float r = sqrtf(c.x * c.x + c.z * c.z);
        float u = eu * W, v = ev * H;
        float flu = floorf(u), flv = floorf(v);
        float mx = 2.0f * W * ((1.5f * spread / r + 4.0e-7f) * 0.15915494f + 6.0e-8f + 1.2e-7f);
        float my = 2.0f * H * (spread + 4.0e-7f + 1.2e-7f);
        bool stable = r >= 1.0e-3f && u - flu > mx && u - flu < 1.0f - mx && v - flv > my && v - flv < 1.0f - my &&
                      flu >= 0.0f && flv >= 0.0f && flu < W && flv < H && mx < 0.25f && my < 0.25f;
        if (!stable) return false; // NaN coordinates end here as well
        v4f p = ld4(img.pixels + (size_t)flv * img.width + (size_t)flu);
        E = mk3(p.x, p.y, p.z);
    }
    const float weight = 1.0f / (float)spp;
    const f3 term = mul3(mk3(E.x + 0.0f, E.y + 0.0f, E.z + 0.0f), weight);
    total = mk3(0.0f, 0.0f, 0.0f);
    for (uint32_t s = 0; s < spp; ++s) total = add3(total, term);
    return true;
}

struct PathCounters { uint32_t rays, hits, misses; };

// One light path = one iteration of the sample loop of sp_PathTraceTile
// (simd_path_tracer.cpp:216-320).  `rng` continues the caller's stream.
template <int MATH, int ENVFILTER, bool CULL, int STEPPED = 0>
SPB_HD f3 trace_path(const DScene &S, const DMaterials &M, const DCamera &cam, uint32_t x,
                     uint32_t y, uint32_t &rng, uint32_t bounceCount, float clampValue,
                     uint32_t *stack, float *stackT, PathCounters &pc, Counters *counters)
{
    f3 o, d;
    primary_ray(cam, x, y, rng, o, d);

    VertexTerms path[SPB_MAX_BOUNCES];
    uint32_t pathLength = 0;
    for (uint32_t bounce = 0; bounce < bounceCount; ++bounce)
    {
        Hit hit = STEPPED == 1 ? intersect_scene_stepped<CULL>(S, o, d, stack, stackT, counters)
                  : STEPPED == 2 ? intersect_scene_stepped2<CULL>(S, o, d, stack, stackT, counters)
                                 : intersect_scene<CULL>(S, o, d, stack, stackT, counters);
        pc.rays++;
        f3 V = neg3(d);
        if (hit.t > 0.0f)
        {
            Surface sf = resolve_hit(S, hit);
            f3 P = add3(o, mul3(d, hit.t));
            f3 L = random_hemisphere<MATH>(sf.normal, rng);
            path[pathLength++] =
                vertex_terms<MATH, ENVFILTER>(M, sf.material, L, sf.normal, V, sf.uvx, sf.uvy, counters);
            o = add3(P, mul3(sf.normal, 0.0001f));
            d = L;
            pc.hits++;
        }
        else
        {
            f3 zero = mk3(0.0f, 0.0f, 0.0f);
            path[pathLength++] =
                vertex_terms<MATH, ENVFILTER>(M, M.backgroundId, zero, zero, V, 0.0f, 0.0f, counters);
            pc.misses++;
            break;
        }
    }
    f3 radiance = mk3(0.0f, 0.0f, 0.0f);
    for (int i = (int)pathLength - 1; i >= 0; --i)
    {
        radiance = fold_radiance(path[i], radiance, clampValue);
    }
    return radiance;
}

} // namespace spb
