// FP64 throughput probe for B200 (sm_100a): DFMA vs DMMA m8n8k4 vs m16n8k8.
// Prints achieved TFLOP/s for each so DESIGN.md can state the measured FP64 peak.
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int ILP>
__global__ void dfma_kernel(double* out, int iters, double a, double b) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void dmma884_kernel(double* out, int iters, double a, double b) {
    double c[ILP][2];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void dmma1688_kernel(double* out, int iters, double a, double b) {
    double c[ILP][4];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F f, int reps) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); f();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0);
        f();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    printf("device %s sms %d clock %d kHz\n", p.name, sms, p.clockRate);
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 16 * 1024));
    const int iters = 20000;
    for (int wpsm : {4, 8, 16, 32}) {
        int threads = 256, blocks = sms * (wpsm * 32 / threads > 0 ? wpsm * 32 / threads : 1);
        if (wpsm * 32 < threads) { threads = wpsm * 32; blocks = sms; }
        {
            float ms = time_ms([&] { dfma_kernel<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
            double fl = 2.0 * 8 * iters * (double)blocks * threads;
            printf("DFMA      warps/SM %2d: %.3f ms  %.2f TFLOP/s\n", wpsm, ms, fl / ms * 1e-9);
        }
        {
            float ms = time_ms([&] { dmma884_kernel<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
            double fl = 2.0 * 256 * 8 * iters * (double)blocks * threads / 32;
            printf("DMMA884   warps/SM %2d: %.3f ms  %.2f TFLOP/s\n", wpsm, ms, fl / ms * 1e-9);
        }
        {
            float ms = time_ms([&] { dmma1688_kernel<4><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
            double fl = 2.0 * 1024 * 4 * iters * (double)blocks * threads / 32;
            printf("DMMA1688  warps/SM %2d: %.3f ms  %.2f TFLOP/s\n", wpsm, ms, fl / ms * 1e-9);
        }
    }
    CK(cudaDeviceSynchronize());
    printf("done\n");
    return 0;
}
