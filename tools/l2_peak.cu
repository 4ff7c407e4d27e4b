// tools/l2_peak.cu -- L2, L1 and HBM read-bandwidth micro-benchmark ("L2 peak must
// be measured by a read micro-benchmark and recorded next to the HBM peak").  Not product code.
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/l2_peak tools/l2_peak.cu
//   tools/l2_peak > gpurun_out/l2_peak.json
//
// Round 1's version swept two working sets with ld.global.nc (which allocates in L1) and reported
// 10.8 TB/s for 16 MiB but 17.3 TB/s for 48 MiB; a smaller set cannot be slower out of L2, so part
// of the larger figure had to come from L1 (several CTAs of an SM re-reading lines a neighbour CTA
// had just pulled in).  This version separates the levels:
//   * "l2"   sweeps: ld.global.cg (cached in L2 only, never in L1), working sets 4 ... 96 MiB;
//   * "l1l2" sweeps: the same addresses with ld.global.nc (L1 allocating), what round 1 measured;
//   * 32-byte loads (ld.global.nc.v8 / LDG.E.256, what the traversal kernel's node fetches are) at
//     the 32 MiB set, both ways;
//   * 64 KiB per CTA re-read in place: the L1 path;
//   * 1 GiB: HBM read-only, for comparison with MEASURED_PEAKS.json's copy figure.
// Best of 5 launches after one warm-up launch, CUDA events.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int MODE> __device__ __forceinline__ float load_sum(const float4 *p)
{
    float4 v;
    if (MODE == 0)
        asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    else
        asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v.x + v.y + v.z + v.w;
}

// MODE 0: .cg (L2 only), 1: .nc (L1 + L2).  Every thread issues `loads` loads whatever the working
// set is: element i, then i + step (mod count), ...; step = grid threads mod count, so consecutive
// loads of a thread are far apart and a CTA never re-reads a line it just had.  (Round 1's version
// looped over passes of count / threads loads with a 64-bit modulo per pass: for small sets that is
// one to three loads per modulo, and the figure measured the index arithmetic, not the cache.)
template <int MODE>
__global__ void __launch_bounds__(256) k_sweep(const float4 *data, unsigned count, unsigned step, int loads, float *sink)
{
    float acc = 0.0f;
    unsigned i = (unsigned)((size_t)(blockIdx.x * blockDim.x + threadIdx.x) % count);
#pragma unroll 8
    for (int k = 0; k < loads; ++k)
    {
        acc += load_sum<MODE>(data + i);
        i += step;
        if (i >= count) i -= count;
    }
    if (acc == 123.456f) *sink = acc;
}

// 32-byte loads (count in 32-byte units)
template <int MODE>
__global__ void __launch_bounds__(256) k_sweep32(const float4 *data, unsigned count, unsigned step, int loads, float *sink)
{
    float acc = 0.0f;
    unsigned i = (unsigned)((size_t)(blockIdx.x * blockDim.x + threadIdx.x) % count);
#pragma unroll 4
    for (int k = 0; k < loads; ++k)
    {
        float4 a, b;
        const float4 *q = data + (size_t)i * 2;
        if (MODE == 0)
            asm volatile("ld.global.cg.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(q));
        else
            asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(q));
        acc += a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w;
        i += step;
        if (i >= count) i -= count;
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
        for (int k = threadIdx.x; k < 4096; k += 256) acc += load_sum<1>(mine + k);
    if (acc == 123.456f) *sink = acc;
}

static cudaEvent_t e0, e1;

template <class F> static double best_of(F launch, double bytes)
{
    double best = 0;
    for (int rep = 0; rep < 6; ++rep)
    {
        CHECK(cudaEventRecord(e0));
        launch();
        CHECK(cudaEventRecord(e1));
        CHECK(cudaEventSynchronize(e1));
        CHECK(cudaGetLastError());
        float ms = 0;
        CHECK(cudaEventElapsedTime(&ms, e0, e1));
        double gbs = bytes / (ms * 1e-3) / 1e9;
        if (rep > 0 && gbs > best) best = gbs;
    }
    return best;
}

int main()
{
    int sms = 0, clockKhz = 0;
    CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    CHECK(cudaDeviceGetAttribute(&clockKhz, cudaDevAttrClockRate, 0));
    const size_t maxBytes = (size_t)1 << 30;
    float4 *data;
    float *sink;
    CHECK(cudaMalloc(&data, maxBytes));
    CHECK(cudaMalloc(&sink, 4));
    CHECK(cudaMemset(data, 0, maxBytes));
    CHECK(cudaEventCreate(&e0));
    CHECK(cudaEventCreate(&e1));
    const int grid = sms * 8;
    const size_t threads = (size_t)grid * 256;
    printf("{\"sms\": %d, \"sm_clock_mhz_max\": %d", sms, clockKhz / 1000);
    const int mib[] = {2, 4, 8, 16, 24, 32, 48, 64, 96};
    const int nsets = 9;
    double l2Best = 0, l1l2Best = 0;
    auto step_for = [&](unsigned count) {
        unsigned step = (unsigned)(threads % count);
        if (step < 4096) step += 4096 * 3 + 1; // never re-read the same line back to back
        return step % count;
    };
    for (int mode = 0; mode < 2; ++mode)
    {
        printf(", \"%s\": {", mode == 0 ? "l2_cg_16B_GBs" : "l1l2_nc_16B_GBs");
        for (int s = 0; s < nsets; ++s)
        {
            unsigned count = (unsigned)(((size_t)mib[s] << 20) / 16);
            const int loads = 2048;
            double bytes = (double)threads * loads * 16;
            unsigned step = step_for(count);
            double g = mode == 0 ? best_of([&] { k_sweep<0><<<grid, 256>>>(data, count, step, loads, sink); }, bytes)
                                 : best_of([&] { k_sweep<1><<<grid, 256>>>(data, count, step, loads, sink); }, bytes);
            printf("%s\"%dMiB\": %.1f", s ? ", " : "", mib[s], g);
            if (mode == 0 && g > l2Best) l2Best = g;
            if (mode == 1 && g > l1l2Best) l1l2Best = g;
        }
        printf("}");
    }
    {
        unsigned count = (unsigned)(((size_t)32 << 20) / 32);
        const int loads = 1024;
        double bytes = (double)threads * loads * 32;
        unsigned step = step_for(count);
        double a = best_of([&] { k_sweep32<0><<<grid, 256>>>(data, count, step, loads, sink); }, bytes);
        double b = best_of([&] { k_sweep32<1><<<grid, 256>>>(data, count, step, loads, sink); }, bytes);
        printf(", \"l2_cg_32B_32MiB_GBs\": %.1f, \"l1l2_nc_32B_32MiB_GBs\": %.1f", a, b);
        if (a > l2Best) l2Best = a;
    }
    {
        unsigned count = (unsigned)(maxBytes / 16);
        const int loads = 1024;
        double bytes = (double)threads * loads * 16;
        double g = best_of([&] { k_sweep<0><<<grid, 256>>>(data, count, step_for(count), loads, sink); }, bytes);
        printf(", \"hbm_1GiB_GBs\": %.1f", g);
    }
    double l1 = best_of([&] { k_l1<<<grid, 256>>>(data, 2000, sink); }, (double)grid * 65536.0 * 2000);
    printf(", \"l1_64KiB_per_cta_GBs\": %.1f", l1);
    // the figures the bench line uses: L2 = best L2-only sweep; the mixed figure is kept for the record
    printf(", \"l2_read_peak_GBs\": %.1f, \"l1l2_mixed_best_GBs\": %.1f", l2Best, l1l2Best);
    printf(", \"how\": \"tools/l2_peak.cu: read sweeps by %d CTAs x 256 threads, best of 5, CUDA events; "
           "l2 = ld.global.cg (L2 only), l1l2 = ld.global.nc (L1 allocating)\"}\n", grid);
    return 0;
}
