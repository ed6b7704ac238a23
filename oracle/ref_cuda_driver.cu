/* TEST INFRASTRUCTURE -- host driver for the reference's own *generated CUDA*.
 *
 * Linked (by oracle/build_ref.py, build_cuda) against the sources that the unmodified
 * reference generator emits with `--lang cuda`, rebuilt for sm_100a.  It does what the
 * reference's performance harness does (pyjac/performance_tester/tester.cu.in:27-168: one
 * thread per state, 64-thread blocks, state-fastest arrays with pitch `padded`, host<->device
 * transfers timed together with the kernel) minus cudaDeviceReset(), and additionally reports
 * the kernel time alone.  Used only by tools/ref_cuda_bench.py as a reported baseline and as a
 * cross-check; never by the product path.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>

#include <cuda.h>
#include <cuda_runtime.h>
#include <helper_cuda.h>
#include "header.cuh"
#include "gpu_memory.cuh"
#include "launch_bounds.cuh"
#include "jacob.cuh"

__global__ void jac_driver(int NUM, const double* pres, const double* y, const mechanism_memory* d_mem)
{
    if (T_ID < NUM) eval_jacob(0, pres[T_ID], y, d_mem->jac, d_mem);
}

static double now_ms(void)
{
    struct timeval tv;
    gettimeofday(&tv, NULL);
    return tv.tv_sec * 1e3 + tv.tv_usec * 1e-3;
}

extern "C" int refcu_nsp(void) { return NSP; }

/* y_sf: NSP x num (T, Y_0..Y_{NSP-2}), state-fastest with pitch num; jac_sf: NSP*NSP x num or NULL.
 * total_ms: H2D + kernel + D2H once, as the reference harness times it; kernel_ms: best of
 * `repeats` launches (CUDA events; jac re-zeroed outside the timed region). */
extern "C" int refcu_run(int num, const double* pres, const double* y_sf, double* jac_sf, int repeats,
                         double* kernel_ms, double* total_ms)
{
    const int padded = (num + TARGET_BLOCK_SIZE - 1) / TARGET_BLOCK_SIZE * TARGET_BLOCK_SIZE;
    size_t free_mem = 0, total_mem = 0;
    if (cudaMemGetInfo(&free_mem, &total_mem) != cudaSuccess) return -1;
    if ((double)required_mechanism_size() * padded > 0.8 * (double)free_mem) return -2;
    mechanism_memory* d_mem = 0;
    mechanism_memory* h_mem = (mechanism_memory*)malloc(sizeof(mechanism_memory));
    initialize_gpu_memory(padded, &h_mem, &d_mem);
    size_t smem = 0;
#ifdef SHARED_SIZE
    smem = SHARED_SIZE;
#endif
    dim3 grid(padded / TARGET_BLOCK_SIZE, 1), block(TARGET_BLOCK_SIZE, 1);
    double* jac_host = jac_sf ? jac_sf : (double*)malloc((size_t)num * NSP * NSP * sizeof(double));

    cudaDeviceSynchronize();
    const double t0 = now_ms();
    cudaMemcpy(h_mem->var, pres, num * sizeof(double), cudaMemcpyHostToDevice);
    cudaMemcpy2D(h_mem->y, padded * sizeof(double), y_sf, num * sizeof(double), num * sizeof(double), NSP,
                 cudaMemcpyHostToDevice);
    jac_driver<<<grid, block, smem>>>(num, h_mem->var, h_mem->y, d_mem);
    cudaMemcpy2D(jac_host, num * sizeof(double), h_mem->jac, padded * sizeof(double), num * sizeof(double),
                 NSP * NSP, cudaMemcpyDeviceToHost);
    cudaDeviceSynchronize();
    *total_ms = now_ms() - t0;
    if (!jac_sf) free(jac_host);
    int rc = cudaGetLastError() == cudaSuccess ? 0 : -3;

    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 1e30;
    for (int r = 0; r < repeats && rc == 0; ++r) {
        cudaMemset(h_mem->jac, 0, (size_t)NSP * NSP * padded * sizeof(double));
        cudaMemset(h_mem->spec_rates, 0, (size_t)NSP * padded * sizeof(double));
        cudaMemset(h_mem->dy, 0, (size_t)NSP * padded * sizeof(double));
        cudaEventRecord(e0);
        jac_driver<<<grid, block, smem>>>(num, h_mem->var, h_mem->y, d_mem);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    *kernel_ms = best;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    free_gpu_memory(&h_mem, &d_mem);
    free(h_mem);
    return rc;
}
