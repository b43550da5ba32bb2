// FP64 tensor-pipe peak probe: the roofline denominator bench.py reports is measured on the box the bench runs on.
// A register-resident chain of DMMA.8x8x4 (mma.sync.m8n8k4.f64) with 8 independent accumulators per warp, 16 warps per SM.
#include <cuda_runtime.h>
#include "common.cuh"
#include "../../include/mspde.h"

namespace {

__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters, double a, double b) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[size_t(blockIdx.x) * blockDim.x + threadIdx.x] = s;
}

}  // namespace

extern "C" int mspde_fp64_peak_probe(void* stream, double* tflops_host) {
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0, sms = 0;
    MSPDE_CUDA(cudaGetDevice(&dev));
    MSPDE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int threads = 256, blocks = sms * 2, iters = 40000;     // 16 warps per SM
    double* out = nullptr;
    MSPDE_CUDA(cudaMalloc(&out, sizeof(double) * size_t(blocks) * threads));
    cudaEvent_t e0, e1;
    MSPDE_CUDA(cudaEventCreate(&e0));
    MSPDE_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {      // first two repetitions are the warm-up; best of the rest
        MSPDE_CUDA(cudaEventRecord(e0, st));
        dmma_peak_kernel<<<blocks, threads, 0, st>>>(out, iters, 1.0000001, 1e-9);
        MSPDE_CUDA(cudaEventRecord(e1, st));
        MSPDE_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        MSPDE_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep >= 2 && ms < best) best = ms;
    }
    MSPDE_LAUNCH_CHECK();
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    const double flops = 2.0 * 256 * 8 * double(iters) * double(blocks) * threads / 32;   // 8x8x4 FMA = 512 flop per DMMA
    *tflops_host = flops / (double(best) * 1e-3) * 1e-12;
    return 0;
}
