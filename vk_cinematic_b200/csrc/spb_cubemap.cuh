// spb_cubemap.cuh -- per-texel arithmetic of the reference's environment pre-processing
// (src/cubemap.cpp): equirectangular map -> cube map faces (CreateCubeMap, :237-291) and the
// diffuse irradiance cube map (CreateIrradianceCubeMap, :108-233), both through
// SampleImageBilinear (src/image.h:34-73).  SURVEY.md §8(f) row 4.
//
// One definition for host and device (like spb_core.cuh): the kernels in spb_cubemap.cu and the
// host build in tests/hostsim call the same functions.  Operation order follows the reference
// line by line; compiled without FMA contraction on both sides.
#pragma once
#include "spb_core.cuh"

namespace spb {

// MapCubeMapLayerIndexToBasisVectors (cubemap.cpp:54-104): layers +X -X +Y -Y +Z -Z
SPB_HD void cube_face_basis(uint32_t layer, f3 &forward, f3 &up, f3 &right)
{
    switch (layer)
    {
    case 0: forward = mk3(1, 0, 0); up = mk3(0, 1, 0); right = mk3(0, 0, -1); break;
    case 1: forward = mk3(-1, 0, 0); up = mk3(0, 1, 0); right = mk3(0, 0, 1); break;
    case 2: forward = mk3(0, 1, 0); up = mk3(0, 0, -1); right = mk3(1, 0, 0); break;
    case 3: forward = mk3(0, -1, 0); up = mk3(0, 0, 1); right = mk3(1, 0, 0); break;
    case 4: forward = mk3(0, 0, 1); up = mk3(0, 1, 0); right = mk3(1, 0, 0); break;
    default: forward = mk3(0, 0, -1); up = mk3(0, 1, 0); right = mk3(-1, 0, 0); break;
    }
}

// texel -> direction (cubemap.cpp:263-275, same lines at :139-150):
// f = x / width, y flipped, mapped to [-1, 1); dir = Normalize(forward + right*fx + up*fy)
SPB_HD f3 cube_texel_direction(f3 forward, f3 up, f3 right, uint32_t x, uint32_t y, uint32_t width,
                               uint32_t height)
{
    float fx = (float)x / (float)width;
    float fy = (float)y / (float)height;
    fy = 1.0f - fy;
    fx = fx * 2.0f - 1.0f;
    fy = fy * 2.0f - 1.0f;
    f3 dir = add3(add3(forward, mul3(right, fx)), mul3(up, fy));
    return normalize3(dir);
}

// SampleImageBilinear (image.h:34-73) with all four channels; Lerp(a, b, t) = a*(1-t) + b*t
// (math_lib.h:432-436).  uv in [0, 1] (the reference asserts it); x0/y0 are clamped as well so
// that a NaN coordinate reads texel 0 instead of faulting.
SPB_HD v4f sample_bilinear4(const DImage &img, float u, float v)
{
    float px = (u * (float)img.width) - 0.5f;
    float py = (v * (float)img.height) - 0.5f;
    px = rmax(px, 0.0f);
    py = rmax(py, 0.0f);
    float flx = floorf(px), fly = floorf(py);
    uint32_t x0 = flx > 0.0f ? (flx < 4294967040.0f ? (uint32_t)flx : 0xFFFFFF00u) : 0u;
    uint32_t y0 = fly > 0.0f ? (fly < 4294967040.0f ? (uint32_t)fly : 0xFFFFFF00u) : 0u;
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
    v4f t0, t1, r;
    t0.x = s0.x * ax + s1.x * fx; t0.y = s0.y * ax + s1.y * fx;
    t0.z = s0.z * ax + s1.z * fx; t0.w = s0.w * ax + s1.w * fx;
    t1.x = s2.x * ax + s3.x * fx; t1.y = s2.y * ax + s3.y * fx;
    t1.z = s2.z * ax + s3.z * fx; t1.w = s2.w * ax + s3.w * fx;
    r.x = t0.x * ay + t1.x * fy; r.y = t0.y * ay + t1.y * fy;
    r.z = t0.z * ay + t1.z * fy; r.w = t0.w * ay + t1.w * fy;
    return r;
}

// direction -> filtered environment texel (cubemap.cpp:277-281, :185-190, :213-216)
template <int MATH>
SPB_HD v4f env_lookup_bilinear(const DImage &env, f3 dir)
{
    float eu, ev;
    equirect_uv<MATH>(dir, eu, ev);
    return sample_bilinear4(env, eu, ev);
}

// One texel of CreateCubeMap (cubemap.cpp:259-287)
template <int MATH>
SPB_HD v4f cube_map_texel(const DImage &env, uint32_t layer, uint32_t x, uint32_t y, uint32_t width,
                          uint32_t height)
{
    f3 forward, up, right;
    cube_face_basis(layer, forward, up, right);
    f3 dir = cube_texel_direction(forward, up, right, x, y, width, height);
    return env_lookup_bilinear<MATH>(env, dir);
}

// Clamp(v, 0, c) = Min(Max(v, 0), c) with the reference's ternaries (math_utils.h:9-19,135-139)
SPB_HD float clamp0(float v, float c) { return rmin(rmax(v, 0.0f), c); }

// Tangent frame of the uniform-sampling branch (cubemap.cpp:156-158)
SPB_HD void irradiance_frame(f3 up, f3 normal, f3 &tangent, f3 &bitangent)
{
    tangent = normalize3(cross3(up, normal));
    bitangent = normalize3(cross3(normal, tangent));
}

// One (phi, theta) sample of the uniform-sampling branch (cubemap.cpp:166-195): the summand
// radiance * Cos(theta) * Sin(theta), products left to right
template <int MATH>
SPB_HD f3 irradiance_uniform_term(const DImage &env, f3 normal, f3 tangent, f3 bitangent, float phi,
                                  float theta, float clampValue)
{
    // MapSphericalToCartesianCoordinates(Vec2(phi, theta)) (math_lib.h:864-871)
    float sinTheta = m_sin<MATH>(theta), cosTheta = m_cos<MATH>(theta);
    float tx = sinTheta * m_cos<MATH>(phi);
    float tz = sinTheta * m_sin<MATH>(phi);
    float ty = cosTheta;
    f3 worldDir = add3(add3(mul3(normal, ty), mul3(tangent, tx)), mul3(bitangent, tz));
    v4f s = env_lookup_bilinear<MATH>(env, worldDir);
    f3 radiance = mk3(clamp0(s.x, clampValue), clamp0(s.y, clampValue), clamp0(s.z, clampValue));
    return mul3(mul3(radiance, cosTheta), sinTheta);
}

// The closing line of the uniform branch (cubemap.cpp:198): PI * irradiance * (1 / count)
SPB_HD f3 irradiance_uniform_finish(f3 sum, uint32_t sampleCount)
{
    return mul3(mul3(sum, SPB_PI), 1.0f / (float)sampleCount);
}

// One sample of the random-sampling branch (cubemap.cpp:201-224), the branch config.h:47 turns
// off.  Vec3(RandomBilateral, RandomBilateral, RandomBilateral): g++ evaluates the three
// arguments right to left, so z is drawn first (same rule as the jitter).
template <int MATH>
SPB_HD f3 irradiance_random_term(const DImage &env, f3 dir, uint32_t &rng, float clampValue,
                                 float sampleContribution)
{
    float oz = rand_bilateral(rng);
    float oy = rand_bilateral(rng);
    float ox = rand_bilateral(rng);
    f3 sampleDir = normalize3(add3(dir, mk3(ox, oy, oz)));
    if (dot3(sampleDir, dir) < 0.0f) sampleDir = neg3(sampleDir);
    float cosine = rmax(dot3(dir, sampleDir), 0.0f);
    v4f s = env_lookup_bilinear<MATH>(env, sampleDir);
    f3 radiance = mk3(clamp0(s.x * cosine, clampValue), clamp0(s.y * cosine, clampValue),
                      clamp0(s.z * cosine, clampValue));
    return mul3(radiance, sampleContribution);
}

// XorShift32 is linear over GF(2): state after k steps = M^k * state.  A jump table holds the
// columns of M^(stride * 2^j), j = 0..31 (32 words each); advancing a state by n * stride steps
// is one matrix-vector product per set bit of n.  That is how texel i of the random-sampling
// branch finds its place in the reference's single serial stream (rng.state = 0x45BA12F3 once,
// cubemap.cpp:122-123, three draws per sample) without walking it.
SPB_HD uint32_t xorshift_apply(const uint32_t *matrixColumns, uint32_t state)
{
    uint32_t r = 0;
    for (int b = 0; b < 32; ++b)
    {
        if ((state >> b) & 1u) r ^= matrixColumns[b];
    }
    return r;
}
SPB_HD uint32_t xorshift_jump(const uint32_t *jumpTable, uint32_t n, uint32_t state)
{
    for (int j = 0; j < 32 && (n >> j) != 0u; ++j)
    {
        if ((n >> j) & 1u) state = xorshift_apply(jumpTable + 32 * j, state);
    }
    return state;
}

// Host side of the jump: columns of M^stride (by walking each basis state stride steps), then
// repeated squaring.  table = 32 levels x 32 words.
inline void xorshift_build_jump_table(uint32_t stride, uint32_t *table)
{
    for (uint32_t b = 0; b < 32; ++b)
    {
        uint32_t state = 1u << b;
        for (uint32_t k = 0; k < stride; ++k) xorshift32(state);
        table[b] = state;
    }
    for (uint32_t j = 1; j < 32; ++j)
        for (uint32_t b = 0; b < 32; ++b)
            table[32 * j + b] = xorshift_apply(table + 32 * (j - 1), table[32 * (j - 1) + b]);
}

// The (phi, theta) values of the reference's float-accumulating loops (cubemap.cpp:162-164):
// start 0, add sampleDelta while below the bound (2*PI resp. 0.5*PI, in float)
inline uint32_t irradiance_loop_values(float bound, float sampleDelta, float *out, uint32_t maxCount)
{
    uint32_t n = 0;
    for (float v = 0.0f; v < bound; v += sampleDelta)
    {
        if (n < maxCount && out) out[n] = v;
        ++n;
    }
    return n;
}

} // namespace spb
