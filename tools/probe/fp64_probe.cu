// FP64 latency / throughput probe for the look-ahead chain of the reduced-system Cholesky (one warp, one SM).
#include <cuda_runtime.h>
#include <cstdio>
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__global__ void k(double* out, long long* clk, double seed, int nwarps_busy) {
    __shared__ double sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = seed + i * 1e-3;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (wid != 0) {
        // optional background load on other warps: independent DFMAs
        if (wid <= nwarps_busy) {
            double x0 = seed, x1 = seed + 1, x2 = seed + 2, x3 = seed + 3;
            for (int i = 0; i < 20000; ++i) { x0 = fma(x0, 1.0000001, 1e-9); x1 = fma(x1, 1.0000001, 1e-9); x2 = fma(x2, 1.0000001, 1e-9); x3 = fma(x3, 1.0000001, 1e-9); }
            out[64 + threadIdx.x] = x0 + x1 + x2 + x3;
        }
        return;
    }
    long long t0, t1;
    double x = seed + lane * 1e-3, y = seed * 0.5;
    const int N = 256;
    // 0: dependent DFMA chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = fma(x, 1.0000001, y);
    t1 = clock64(); if (lane == 0) clk[0] = (t1 - t0);
    // 1: dependent DMUL chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = x * 1.0000001;
    t1 = clock64(); if (lane == 0) clk[1] = (t1 - t0);
    // 2: dependent rsqrt chain
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; ++i) x = rsqrt(x + 2.0);
    t1 = clock64(); if (lane == 0) clk[2] = (t1 - t0);
    // 3: dependent shfl (64-bit) chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = __shfl_sync(0xffffffffu, x, (i + 1) & 31);
    t1 = clock64(); if (lane == 0) clk[3] = (t1 - t0);
    // 4: independent DFMA throughput, 8 chains
    double c[8];
    for (int j = 0; j < 8; ++j) c[j] = x + j;
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) c[j] = fma(c[j], 1.0000001, y);
    t1 = clock64(); if (lane == 0) clk[4] = (t1 - t0);
    for (int j = 0; j < 8; ++j) x += c[j];
    // 5: dependent DMMA chain (same accumulator)
    double d0 = x, d1 = y;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) dmma(d0, d1, 1.0000001, 0.25);
    t1 = clock64(); if (lane == 0) clk[5] = (t1 - t0);
    // 6: independent DMMAs, 8 accumulators
    double e[8][2];
    for (int j = 0; j < 8; ++j) { e[j][0] = d0 + j; e[j][1] = d1; }
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) dmma(e[j][0], e[j][1], 1.0000001, 0.25);
    t1 = clock64(); if (lane == 0) clk[6] = (t1 - t0);
    for (int j = 0; j < 8; ++j) x += e[j][0] + e[j][1];
    // 7: dependent LDS chain (pointer chasing by value)
    int idx = lane;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) idx = ((int)sm[idx & 1023] + idx + 1) & 1023;
    t1 = clock64(); if (lane == 0) clk[7] = (t1 - t0);
    // 8: LDS (broadcast) + dependent DFMA: the substitution inner pattern, 1 chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = fma(x, sm[(i * 7) & 1023], y);
    t1 = clock64(); if (lane == 0) clk[8] = (t1 - t0);
    // 9: full 32-row substitution as written in the kernel (32 columns), LdT in shared memory
    {
        double xr[32];
#pragma unroll
        for (int cidx = 0; cidx < 32; ++cidx) xr[cidx] = x + cidx + lane;
        t0 = clock64();
#pragma unroll
        for (int cidx = 0; cidx < 32; ++cidx) {
            const double xc = xr[cidx] * sm[992 + cidx];
            xr[cidx] = xc;
#pragma unroll
            for (int c2 = cidx + 1; c2 < 32; ++c2) xr[c2] = fma(-xc, sm[(cidx * 32 + c2) & 1023], xr[c2]);
        }
        t1 = clock64(); if (lane == 0) clk[9] = (t1 - t0);
#pragma unroll
        for (int cidx = 0; cidx < 32; ++cidx) x += xr[cidx];
    }
    out[threadIdx.x] = x + idx + d0 + d1;
}
int main() {
    double* out; long long* clk;
    cudaMalloc(&out, 8192); cudaMalloc(&clk, 128);
    const char* names[10] = {"dep DFMA (per op)", "dep DMUL (per op)", "dep rsqrt(double) (per op)", "dep shfl 64-bit (per op)", "indep DFMA x8 (per op)",
                             "dep DMMA (per op)", "indep DMMA x8 (per op)", "dep LDS+cvt chain (per op)", "LDS-fed dep DFMA (per op)", "32-col substitution (total cycles)"};
    for (int busy = 0; busy <= 3; busy += 3) {
        for (int rep = 0; rep < 2; ++rep) k<<<1, 128, 0>>>(out, clk, 1.25, busy);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        long long h[16]; cudaMemcpy(h, clk, 128, cudaMemcpyDeviceToHost);
        printf("background busy warps (other schedulers): %d\n", busy);
        for (int i = 0; i < 10; ++i) {
            double per = (i == 9) ? (double)h[i] : (i == 4 || i == 6) ? h[i] / (256.0 * 8) : h[i] / 256.0;
            printf("  %-36s %10.1f\n", names[i], per);
        }
    }
    return 0;
}
