// Dependent-issue latencies on this GPU, in SM cycles (clock64 around a chain of N dependent operations, one warp per block):
// what a warp of k_eval waits for between two instructions when nothing else hides it.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
constexpr int N = 4096;

__global__ void k(double* out, long long* cyc, const int* chase, const uint4* tab)
{
    __shared__ __align__(16) int sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = ((i * 17 + 16) & 4095) & ~3;   // 16-byte aligned hops
    __syncthreads();
    double a = out[0], b = 1.0000001, c = 1e-9;
    long long t0, t1;
    // DFMA chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) a = fma(a, b, c);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    // DADD chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) a = a + c;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[1] = t1 - t0;
    // DMUL chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) a = a * b;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[2] = t1 - t0;
    // shared-memory pointer chase, 128-bit loads
    int p = (threadIdx.x * 4) & 4095;
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; ++i) { const int4 v = *reinterpret_cast<const int4*>(&sm[p]); p = v.x; }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[3] = t1 - t0;
    // shuffle chain (64-bit value = two shuffles)
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; ++i) a = __shfl_sync(0xffffffffu, a, (threadIdx.x + 1) & 31);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[4] = t1 - t0;
    // global pointer chase through L2 (read-only path), 16-byte loads, 1 MB table
    int q = threadIdx.x;
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N / 8; ++i) { const uint4 v = __ldg(tab + q); q = (int)v.x; }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[5] = t1 - t0;
    // int -> address -> shared load -> fp64 use, the usual k_eval step
    p = (threadIdx.x * 4) & 4095;
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; ++i) {
        const int4 v = *reinterpret_cast<const int4*>(&sm[p]);
        a = fma(a, b, (double)v.y * 1e-30);
        p = (v.x + (__double2loint(a) & 0)) & 4095;
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[6] = t1 - t0;
    out[threadIdx.x] = a + p + q;
}

int main()
{
    double* out; long long* cyc; int* chase = nullptr; uint4* tab;
    const int words = 65536;                                   // 1 MB
    CK(cudaMalloc(&out, 1024 * 8)); CK(cudaMemset(out, 0, 1024 * 8));
    CK(cudaMalloc(&cyc, 64)); CK(cudaMalloc(&tab, words * 16));
    uint4* h = new uint4[words];
    for (int i = 0; i < words; ++i) h[i] = uint4{(unsigned)((i * 4099 + 977) % words), 0, 0, 0};
    CK(cudaMemcpy(tab, h, words * 16, cudaMemcpyHostToDevice));
    for (int rep = 0; rep < 2; ++rep) { k<<<1, 32>>>(out, cyc, chase, tab); CK(cudaDeviceSynchronize()); }
    long long c[8]; CK(cudaMemcpy(c, cyc, 56, cudaMemcpyDeviceToHost));
    const char* names[7] = {"DFMA -> DFMA", "DADD -> DADD", "DMUL -> DMUL", "LDS.128 -> address -> LDS.128", "SHFL (64-bit: two) -> SHFL",
                            "LDG.128 (L2 hit) -> address -> LDG.128", "LDS.128 -> I2F / DFMA -> address -> LDS.128"};
    const int cnt[7] = {N, N, N, N, N, N / 8, N};
    printf("| dependent chain | cycles per step |\n|---|---|\n");
    for (int i = 0; i < 7; ++i) printf("| %s | %.1f |\n", names[i], (double)c[i] / cnt[i]);
    return 0;
}
