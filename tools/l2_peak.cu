// tools/l2_peak.cu -- L2 and L1 read-bandwidth micro-benchmark ("L2 peak must be
// measured by a read micro-benchmark and recorded next to the HBM peak").  Not product code.
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/l2_peak tools/l2_peak.cu
//   tools/l2_peak > gpurun_out/l2_peak.json
//
// Every thread streams 16-byte loads (ld.global.nc.v4, what the traversal kernel's node fetches
// are) over a working set that is swept repeatedly: 16 MiB and 48 MiB stay L2-resident on a B200
// (126 MB L2), 1 GiB does not (HBM, for comparison with MEASURED_PEAKS.json), and 64 KiB per CTA
// re-read in place measures the L1 path.  Best of 5 launches, CUDA events.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void __launch_bounds__(256) k_sweep(const float4 *data, size_t count, int passes, float *sink)
{
    float acc = 0.0f;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int p = 0; p < passes; ++p)
    {
        // rotate the start per pass so that a CTA does not re-read the lines it just had in L1
        size_t start = ((size_t)blockIdx.x * blockDim.x + threadIdx.x + (size_t)p * 977 * blockDim.x) % count;
        size_t n = count / stride;
        size_t i = start;
#pragma unroll 8
        for (size_t k = 0; k < n; ++k)
        {
            float4 v = __ldg(data + i);
            acc += v.x + v.y + v.z + v.w;
            i += stride;
            if (i >= count) i -= count;
        }
    }
    if (acc == 123.456f) *sink = acc;
}

// each CTA re-reads its own 64 KiB: L1-resident after the first pass
__global__ void __launch_bounds__(256) k_l1(const float4 *data, int passes, float *sink)
{
    const float4 *mine = data + (size_t)blockIdx.x * 4096;
    float acc = 0.0f;
    for (int p = 0; p < passes; ++p)
#pragma unroll 8
        for (int k = threadIdx.x; k < 4096; k += 256)
        {
            float4 v = __ldg(mine + k);
            acc += v.x + v.y + v.z + v.w;
        }
    if (acc == 123.456f) *sink = acc;
}

int main()
{
    int sms = 0;
    CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const size_t maxBytes = (size_t)1 << 30;
    float4 *data;
    float *sink;
    CHECK(cudaMalloc(&data, maxBytes));
    CHECK(cudaMalloc(&sink, 4));
    CHECK(cudaMemset(data, 0, maxBytes));
    cudaEvent_t e0, e1;
    CHECK(cudaEventCreate(&e0));
    CHECK(cudaEventCreate(&e1));
    const int grid = sms * 8;
    printf("{\"sms\": %d", sms);
    const size_t sets[3] = {(size_t)16 << 20, (size_t)48 << 20, (size_t)1 << 30};
    const char *names[3] = {"l2_16MiB_GBs", "l2_48MiB_GBs", "hbm_1GiB_GBs"};
    for (int s = 0; s < 3; ++s)
    {
        size_t count = sets[s] / 16;
        int passes = s == 2 ? 4 : 64;
        double best = 0;
        for (int rep = 0; rep < 6; ++rep)
        {
            CHECK(cudaEventRecord(e0));
            k_sweep<<<grid, 256>>>(data, count, passes, sink);
            CHECK(cudaEventRecord(e1));
            CHECK(cudaEventSynchronize(e1));
            float ms = 0;
            CHECK(cudaEventElapsedTime(&ms, e0, e1));
            size_t perPass = (count / ((size_t)grid * 256)) * (size_t)grid * 256 * 16;
            double gbs = (double)perPass * passes / (ms * 1e-3) / 1e9;
            if (rep > 0 && gbs > best) best = gbs;
        }
        printf(", \"%s\": %.1f", names[s], best);
    }
    {
        double best = 0;
        const int passes = 2000;
        for (int rep = 0; rep < 6; ++rep)
        {
            CHECK(cudaEventRecord(e0));
            k_l1<<<grid, 256>>>(data, passes, sink);
            CHECK(cudaEventRecord(e1));
            CHECK(cudaEventSynchronize(e1));
            float ms = 0;
            CHECK(cudaEventElapsedTime(&ms, e0, e1));
            double gbs = (double)grid * 65536.0 * passes / (ms * 1e-3) / 1e9;
            if (rep > 0 && gbs > best) best = gbs;
        }
        printf(", \"l1_64KiB_per_cta_GBs\": %.1f", best);
    }
    printf(", \"how\": \"tools/l2_peak.cu: 16-byte __ldg sweeps, %d CTAs x 256 threads, best of 5, CUDA events\"}\n", grid);
    return 0;
}
