// Micro-benchmark: cost of table-word loads through the LSU pipe (ld.global.nc.v4) against the texture pipe
// (tex1Dfetch<int4>), alone and next to a shared-memory load stream like k_eval's (LDS.128, 8 rows of 64 bytes).
// Access pattern of the tables: 8 sub-groups of 4 lanes, each sub-group reads its own 16-byte word, consecutive words.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int MODE>   // 0: ldg only, 1: tex only, 2: lds only, 3: ldg + lds, 4: tex + lds
__global__ void __launch_bounds__(384, 1) k(const int4* __restrict__ tab, cudaTextureObject_t tex, int words, int iters, int* out)
{
    extern __shared__ __align__(16) double sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane >> 2;
    for (int i = threadIdx.x; i < 24 * 1024; i += blockDim.x) sm[i] = i;
    __syncthreads();
    int acc = 0;
    double facc = 0.0;
    int pos = (warp * 997 + blockIdx.x * 131) % (words - 64);
    unsigned row = (warp * 37 + sub * 11) % 2800;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (MODE == 0 || MODE == 3) { const int4 q = __ldg(tab + pos + u * 8 + sub); acc ^= q.x ^ q.y ^ q.z ^ q.w; }
            if (MODE == 1 || MODE == 4) { const int4 q = tex1Dfetch<int4>(tex, pos + u * 8 + sub); acc ^= q.x ^ q.y ^ q.z ^ q.w; }
            if (MODE >= 2) {
#pragma unroll
                for (int g = 0; g < 3; ++g) {     // k_eval does ~2 LDS.128 per table word
                    const double2 v = *reinterpret_cast<const double2*>(reinterpret_cast<const char*>(sm) + row * 64 + (lane & 3) * 16);
                    facc += v.x + v.y;
                    row = (row * 5 + 17 + (unsigned)g) % 2800;
                }
            }
        }
        pos += 32;
        if (pos >= words - 64) pos -= words - 64;
    }
    if (acc == 0x12345678 || facc == 1.2345) out[0] = acc;
}

int main()
{
    const int words = 12800;                 // 200 KB of table words, as the GRI-sized plan
    int4* tab; int* out;
    CK(cudaMalloc(&tab, words * sizeof(int4)));
    CK(cudaMemset(tab, 1, words * sizeof(int4)));
    CK(cudaMalloc(&out, 4));
    cudaResourceDesc rd{}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = tab;
    rd.res.linear.desc = cudaCreateChannelDesc<int4>(); rd.res.linear.sizeInBytes = words * sizeof(int4);
    cudaTextureDesc td{}; td.readMode = cudaReadModeElementType;
    cudaTextureObject_t tex; CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
    const int iters = 4000, smem = 24 * 1024 * 8;
    const char* names[5] = {"ld.global.nc.v4 only", "tex1Dfetch<int4> only", "LDS.128 only (3 per table word)", "ld.global + LDS", "tex + LDS"};
    void (*fn[5])(const int4*, cudaTextureObject_t, int, int, int*) = {k<0>, k<1>, k<2>, k<3>, k<4>};
    for (int m = 0; m < 5; ++m) {
        CK(cudaFuncSetAttribute(fn[m], cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        fn[m]<<<148, 384, smem>>>(tab, tex, words, 100, out);
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        fn[m]<<<148, 384, smem>>>(tab, tex, words, iters, out);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double warp_words = 12.0 * iters * 4;            // table-word loads per SM (warp instructions)
        printf("%-34s %8.3f ms   %.1f cycles per warp-level table load (at 1.965 GHz, 12 warps per SM)\n", names[m], ms, ms * 1e-3 * 1.965e9 / warp_words);
    }
    return 0;
}
