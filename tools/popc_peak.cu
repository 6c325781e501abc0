// popc_peak.cu — micro-benchmark of the integer-pipe peaks the matching roofline is quoted against:
// POPC, LOP3 and IADD3 thread-level throughput of the whole GPU.  Prints one JSON line with operations per second; the SM clock
// is NOT derived here (round 1 divided block 0's clock64 span by the whole kernel's event time and reported 259 MHz): the
// driver script tools/popc_peak.py samples clocks.sm with nvidia-smi while this binary runs and computes the per-clock figures.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/popc_peak tools/popc_peak.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

template <int MODE>
__global__ void __launch_bounds__(256) k(unsigned* out, int iters, unsigned seed, long long* cyc) {
    unsigned v[8], a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[i] = seed * (threadIdx.x + 1) + i * 977u + blockIdx.x; a[i] = 0; }
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { a[i] += __popc(v[i]); v[i] ^= a[i]; }                 // POPC + IADD + LOP3
            if (MODE == 1) { a[i] = (a[i] ^ v[i]) & (v[i] | seed); v[i] ^= a[i]; } // LOP3 only
            if (MODE == 2) { a[i] += v[i]; v[i] += a[i] + seed; }                  // IADD3 only
        }
    }
    const long long t1 = clock64();
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i] ^ v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
double run(int sms, int iters, unsigned* d, long long* dc, double* mhz) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = sms * 8;
    k<MODE><<<grid, 256>>>(d, iters, 12345u, dc);
    cudaDeviceSynchronize();
    float best = 1e30f;
    long long cyc = 0;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        k<MODE><<<grid, 256>>>(d, iters, 12345u + r, dc);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) { best = ms; cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost); }
    }
    *mhz = 0.0;
    (void)cyc;
    return (double)grid * 256 * 8.0 * iters / (best * 1e-3);   // primary ops per second
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    unsigned* d; long long* dc;
    cudaMalloc(&d, (size_t)sms * 8 * 256 * 4); cudaMalloc(&dc, 8);
    double mhz0, mhz1, mhz2;
    double popc = 0, lop = 0, iadd = 0;
    for (int rep = 0; rep < 6; ++rep) {   // ~2 s under load so that nvidia-smi sees the clock this ran at
        popc = run<0>(sms, 1 << 16, d, dc, &mhz0);
        lop = run<1>(sms, 1 << 16, d, dc, &mhz1) * 2;   // two LOP3 per step
        iadd = run<2>(sms, 1 << 16, d, dc, &mhz2) * 2;  // two IADD3 per step
    }
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"popc_per_s\": %.4g, \"lop3_per_s\": %.4g, \"iadd3_per_s\": %.4g, \"sm_clock_attr_mhz\": %.0f}\n",
           p.name, sms, popc, lop, iadd, clk_khz / 1e3);
    return 0;
}
