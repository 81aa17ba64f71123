// sr_diag.cu -- FMA-issue peak microbenchmark (softray_measure_fma_peak): the measured denominator
// of the FP-issue roofline.  Eight independent FMA chains per thread keep the pipe full; the value
// written at the end keeps the compiler from removing the loop.
#include <cuda_runtime.h>

namespace sr {

template <typename T>
__global__ void __launch_bounds__(256) fma_chain_kernel(T* out, int iters, T a, T b)
{
    T x0 = (T)threadIdx.x, x1 = x0 + (T)1, x2 = x0 + (T)2, x3 = x0 + (T)3, x4 = x0 + (T)4, x5 = x0 + (T)5, x6 = x0 + (T)6,
      x7 = x0 + (T)7;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            x0 = x0 * a + b; x1 = x1 * a + b; x2 = x2 * a + b; x3 = x3 * a + b;
            x4 = x4 * a + b; x5 = x5 * a + b; x6 = x6 * a + b; x7 = x7 * a + b;
        }
    }
    const T s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == (T)-12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// returns TFLOP/s (FMA = 2 flops), best of `reps`
cudaError_t measure_fma_peak(bool fp64, int sm_count, cudaStream_t stream, double* tflops)
{
    const int blocks = sm_count * 8, threads = 256, iters = fp64 ? 2048 : 8192, reps = 5;
    void* d = nullptr;
    cudaError_t e = cudaMalloc(&d, (size_t)blocks * threads * sizeof(double));
    if (e != cudaSuccess) return e;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0.0;
    for (int r = 0; r < reps + 1; r++) {
        cudaEventRecord(e0, stream);
        if (fp64) fma_chain_kernel<double><<<blocks, threads, 0, stream>>>((double*)d, iters, 1.0000001, 1e-9);
        else fma_chain_kernel<float><<<blocks, threads, 0, stream>>>((float*)d, iters, 1.0000001f, 1e-9f);
        cudaEventRecord(e1, stream);
        e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 64.0 * (double)iters * (double)blocks * (double)threads;
        if (r > 0 && ms > 0.f) { const double t = flops / (ms * 1e-3) * 1e-12; if (t > best) best = t; }
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d);
    if (e == cudaSuccess) e = cudaGetLastError();
    *tflops = best;
    return e;
}

}  // namespace sr
