// spb_cubemap.cu -- sm_100a kernels of the environment pre-processing step that feeds the path
// (src/cubemap.cpp): equirectangular map -> six cube faces, and the
// diffuse irradiance cube map.  Arithmetic in spb_cubemap.cuh (one definition, host and device).
//
// Compile: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -lineinfo.
//
// k_cube_map: one texel per thread, x fastest, so a warp writes 512 contiguous bytes and its
//   four bilinear taps fall on a few neighbouring rows of the source (a 1024^2 face covers the
//   4096x2048 source at about one texel per texel): HBM-bound, 16 B written per texel and the
//   128 MiB source read about once (algorithmic bytes per texel: 16 + 16 * srcTexels / dstTexels).
// k_irradiance: one CTA per texel.  The reference sums ~1000 terms per texel in a fixed order
//   (float addition does not commute with regrouping), so the terms -- each a handful of
//   double-precision libm calls plus four 16-byte gathers -- are evaluated by all threads into
//   shared memory and then folded front to back by one thread per colour channel: the expensive
//   part runs SPB_IRR_THREADS wide, the order-dependent part stays the reference's.
#include "spb_cubemap.cuh"
#include "spb_kernels.cuh"

namespace spb {

template <int MATH>
__global__ void __launch_bounds__(256)
k_cube_map(DImage env, v4f *out, uint32_t width, uint32_t height)
{
    const uint64_t perFace = (uint64_t)width * height, total = 6 * perFace;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (uint64_t)gridDim.x * blockDim.x)
    {
        uint32_t layer = (uint32_t)(i / perFace);
        uint32_t r = (uint32_t)(i - layer * perFace);
        uint32_t y = r / width, x = r - y * width;
        out[i] = cube_map_texel<MATH>(env, layer, x, y, width, height);
    }
}

#define SPB_IRR_THREADS 256
#define SPB_IRR_CHUNK 1024

// MODE 0: uniform (phi, theta) grid (cubemap.cpp:152-198; config.h:47 selects it);
// MODE 1: random offsets from the reference's single serial XorShift32 stream (cubemap.cpp:200-224)
template <int MATH, int MODE>
__global__ void __launch_bounds__(SPB_IRR_THREADS)
k_irradiance(IrradianceArgs a)
{
    __shared__ float term[3][SPB_IRR_CHUNK];
    const uint32_t perFace = a.width * a.height, total = 6 * perFace;
    const uint32_t sampleCount = MODE == 0 ? a.phiCount * a.thetaCount : a.samplesPerPixel;
    for (uint32_t texel = blockIdx.x; texel < total; texel += gridDim.x)
    {
        uint32_t layer = texel / perFace;
        uint32_t r = texel - layer * perFace;
        uint32_t y = r / a.width, x = r - y * a.width;
        f3 forward, up, right;
        cube_face_basis(layer, forward, up, right);
        f3 dir = cube_texel_direction(forward, up, right, x, y, a.width, a.height);
        f3 tangent = mk3(0, 0, 0), bitangent = mk3(0, 0, 0);
        uint32_t texelState = 0;
        if (MODE == 0) irradiance_frame(up, dir, tangent, bitangent);
        else texelState = xorshift_jump(a.jumpTexel, texel, a.seed);

        float acc = 0.0f; // running sum of channel threadIdx.x / 32 (threads 0, 32, 64)
        for (uint32_t base = 0; base < sampleCount; base += SPB_IRR_CHUNK)
        {
            uint32_t end = sampleCount - base < SPB_IRR_CHUNK ? sampleCount - base : SPB_IRR_CHUNK;
            for (uint32_t k = threadIdx.x; k < end; k += SPB_IRR_THREADS)
            {
                uint32_t s = base + k;
                f3 t;
                if (MODE == 0)
                {
                    uint32_t iphi = s / a.thetaCount, itheta = s - iphi * a.thetaCount;
                    t = irradiance_uniform_term<MATH>(a.env, dir, tangent, bitangent, a.phis[iphi],
                                                      a.thetas[itheta], a.clampValue);
                }
                else
                {
                    uint32_t rng = xorshift_jump(a.jumpSample, s, texelState);
                    t = irradiance_random_term<MATH>(a.env, dir, rng, a.clampValue, a.sampleContribution);
                }
                term[0][k] = t.x;
                term[1][k] = t.y;
                term[2][k] = t.z;
            }
            __syncthreads();
            if ((threadIdx.x & 31u) == 0 && threadIdx.x < 96)
            {
                const float *mine = term[threadIdx.x >> 5];
                for (uint32_t k = 0; k < end; ++k) acc = acc + mine[k];
            }
            __syncthreads();
        }
        if ((threadIdx.x & 31u) == 0 && threadIdx.x < 96)
        {
            float v = acc;
            if (MODE == 0) v = (acc * SPB_PI) * (1.0f / (float)sampleCount); // cubemap.cpp:198
            ((float *)(a.out + texel))[threadIdx.x >> 5] = v;
        }
        if (threadIdx.x == 96) ((float *)(a.out + texel))[3] = 1.0f; // Vec4(irradiance, 1)
    }
}

static unsigned sm_count()
{
    int device = 0, sms = 0;
    cudaGetDevice(&device);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    return sms > 0 ? (unsigned)sms : 1u;
}

void launch_cube_map(const KernelConfig &cfg, const DImage &env, v4f *out, uint32_t width, uint32_t height,
                     cudaStream_t stream)
{
    uint64_t total = 6ull * width * height;
    if (total == 0) return;
    g_kernelLaunches++;
    uint64_t want = (total + 255) / 256, cap = (uint64_t)sm_count() * 8;
    unsigned grid = (unsigned)(want < cap ? want : cap);
    if (cfg.math) k_cube_map<1><<<grid, 256, 0, stream>>>(env, out, width, height);
    else k_cube_map<0><<<grid, 256, 0, stream>>>(env, out, width, height);
}

void launch_irradiance(const KernelConfig &cfg, const IrradianceArgs &args, int mode, cudaStream_t stream)
{
    uint64_t total = 6ull * args.width * args.height;
    if (total == 0) return;
    g_kernelLaunches++;
    uint64_t cap = (uint64_t)sm_count() * 8;
    unsigned grid = (unsigned)(total < cap ? total : cap);
    int key = (cfg.math ? 2 : 0) | (mode ? 1 : 0);
    switch (key)
    {
    case 0: k_irradiance<0, 0><<<grid, SPB_IRR_THREADS, 0, stream>>>(args); break;
    case 1: k_irradiance<0, 1><<<grid, SPB_IRR_THREADS, 0, stream>>>(args); break;
    case 2: k_irradiance<1, 0><<<grid, SPB_IRR_THREADS, 0, stream>>>(args); break;
    default: k_irradiance<1, 1><<<grid, SPB_IRR_THREADS, 0, stream>>>(args); break;
    }
}

} // namespace spb
