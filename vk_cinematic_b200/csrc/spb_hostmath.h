// spb_hostmath.h -- host-side setup arithmetic (camera basis, model matrices, AABB transform).
//
// These run once per scene/camera on the host, exactly as in the reference, and feed values to
// the device that must be bit-identical to what the reference computes; each function follows
// the operation order of the cited reference lines.  Compile with -ffp-contract=off.
#pragma once

#include <math.h>
#include <float.h>
#include <stdint.h>

namespace spbh {

struct V3 { float x, y, z; };
struct V4 { float x, y, z, w; };
struct M4 { V4 c[4]; }; // column-major like the reference's mat4 (math_lib.h:91-95)

inline V3 v3(float x, float y, float z) { V3 r = {x, y, z}; return r; }
inline V3 add(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 sub(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 mul(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
inline V3 neg(V3 a) { return v3(-a.x, -a.y, -a.z); }
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b)
{
    return v3((a.y * b.z) - (a.z * b.y), (a.z * b.x) - (a.x * b.z), (a.x * b.y) - (a.y * b.x));
}
inline float fminr(float a, float b) { return a < b ? a : b; }
inline float fmaxr(float a, float b) { return a > b ? a : b; }

inline float col(const V4 &c, int i) { return i == 0 ? c.x : i == 1 ? c.y : i == 2 ? c.z : c.w; }
inline void setcol(V4 &c, int i, float v)
{
    if (i == 0) c.x = v; else if (i == 1) c.y = v; else if (i == 2) c.z = v; else c.w = v;
}
inline float dot4(V4 a, V4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }

inline M4 identity()
{
    M4 m = {};
    m.c[0].x = 1.0f; m.c[1].y = 1.0f; m.c[2].z = 1.0f; m.c[3].w = 1.0f;
    return m;
}
// math_lib.h:205-226
inline M4 scale(V3 s)
{
    M4 m = {};
    m.c[0].x = s.x; m.c[1].y = s.y; m.c[2].z = s.z; m.c[3].w = 1.0f;
    return m;
}
inline M4 translate(V3 t)
{
    M4 m = identity();
    m.c[3].x = t.x; m.c[3].y = t.y; m.c[3].z = t.z;
    return m;
}
// math_lib.h:250-272: element (row i, column j) = Dot(row i of a, column j of b)
inline M4 matmul(const M4 &a, const M4 &b)
{
    M4 r = {};
    for (int i = 0; i < 4; ++i)
    {
        V4 row = {col(a.c[0], i), col(a.c[1], i), col(a.c[2], i), col(a.c[3], i)};
        for (int j = 0; j < 4; ++j) setcol(r.c[j], i, dot4(row, b.c[j]));
    }
    return r;
}
// math_lib.h:381-398
inline V4 matvec(const M4 &a, V4 b)
{
    V4 r = {};
    for (int i = 0; i < 4; ++i)
    {
        V4 row = {col(a.c[0], i), col(a.c[1], i), col(a.c[2], i), col(a.c[3], i)};
        setcol(r, i, dot4(row, b));
    }
    return r;
}
inline V3 transform_point(V3 p, const M4 &m)
{
    V4 v = matvec(m, V4{p.x, p.y, p.z, 1.0f});
    return v3(v.x, v.y, v.z);
}
// math_lib.h:623-650
inline M4 rotate(V4 q)
{
    M4 m = identity();
    float x = q.x, y = q.y, z = q.z, w = q.w;
    m.c[0].x = 1.0f - 2.0f * y * y - 2.0f * z * z;
    m.c[0].y = 2.0f * x * y + 2.0f * z * w;
    m.c[0].z = 2.0f * x * z - 2.0f * y * w;
    m.c[1].x = 2.0f * x * y - 2.0f * z * w;
    m.c[1].y = 1.0f - 2.0f * x * x - 2.0f * z * z;
    m.c[1].z = 2.0f * y * z + 2.0f * x * w;
    m.c[2].x = 2.0f * x * z + 2.0f * y * w;
    m.c[2].y = 2.0f * y * z - 2.0f * x * w;
    m.c[2].z = 1.0f - 2.0f * x * x - 2.0f * y * y;
    return m;
}
inline V4 conjugate(V4 q) { return V4{-q.x, -q.y, -q.z, q.w}; }
// quat product (math_lib.h:580-586): v = p.s*q.v + q.s*p.v + Cross(p.v,q.v); s = p.s*q.s - Dot
inline V4 quatmul(V4 p, V4 q)
{
    V3 pv = v3(p.x, p.y, p.z), qv = v3(q.x, q.y, q.z);
    V3 v = add(add(mul(qv, p.w), mul(pv, q.w)), cross(pv, qv));
    float s = p.w * q.w - dot(pv, qv);
    return V4{v.x, v.y, v.z, s};
}
// math_lib.h:608-613
inline V3 rotate_vector(V3 v, V4 p)
{
    V4 q = {v.x, v.y, v.z, 0.0f};
    V4 r = quatmul(quatmul(p, q), conjugate(p));
    return v3(r.x, r.y, r.z);
}

// sp_scene.cpp:91-96
inline M4 model_matrix(V3 position, V4 orientation, V3 s)
{
    return matmul(matmul(translate(position), rotate(orientation)), scale(s));
}
inline M4 inverse_model_matrix(V3 position, V4 orientation, V3 s)
{
    V3 invScale = v3(1.0f / s.x, 1.0f / s.y, 1.0f / s.z); // Inverse(), math_lib.h:911-915
    return matmul(matmul(scale(invScale), rotate(conjugate(orientation))), translate(neg(position)));
}

// aabb.h:29-58
inline void transform_aabb(V3 boxMin, V3 boxMax, V3 position, V4 orientation, V3 s, V3 *outMin,
                           V3 *outMax)
{
    M4 m = model_matrix(position, orientation, s);
    V3 corners[8] = {
        v3(boxMin.x, boxMin.y, boxMin.z), v3(boxMax.x, boxMin.y, boxMin.z),
        v3(boxMax.x, boxMin.y, boxMax.z), v3(boxMin.x, boxMin.y, boxMax.z),
        v3(boxMin.x, boxMax.y, boxMin.z), v3(boxMax.x, boxMax.y, boxMin.z),
        v3(boxMax.x, boxMax.y, boxMax.z), v3(boxMin.x, boxMax.y, boxMax.z)};
    V3 lo = transform_point(corners[0], m), hi = lo;
    for (int i = 1; i < 8; ++i)
    {
        V3 p = transform_point(corners[i], m);
        lo = v3(fminr(lo.x, p.x), fminr(lo.y, p.y), fminr(lo.z, p.z));
        hi = v3(fmaxr(hi.x, p.x), fmaxr(hi.y, p.y), fmaxr(hi.z, p.z));
    }
    *outMin = lo;
    *outMax = hi;
}

} // namespace spbh
