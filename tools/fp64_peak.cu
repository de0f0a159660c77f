// fp64_peak.cu -- measures the DFMA issue rate of the device (the fp64 roofline denominator, which
// MEASURED_PEAKS.json does not carry). extern "C" double measure_dfma_tflops(int device, int *sm_count).
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) dfma_kernel(double *out, double a, double b, int iterations) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iterations; i++) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

extern "C" __attribute__((visibility("default"))) double measure_dfma_tflops(int device, int *sm_count) {
    cudaDeviceProp prop;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1.0;
    if (sm_count) *sm_count = prop.multiProcessorCount;
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iterations = 1 << 14;
    double *out = nullptr;
    if (cudaMalloc(&out, sizeof(double) * blocks * threads) != cudaSuccess) return -1.0;
    cudaEvent_t start, stop;
    cudaEventCreate(&start);
    cudaEventCreate(&stop);
    double best = 0.0;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(start);
        dfma_kernel<<<blocks, threads>>>(out, 0.999999, 1.0e-6, iterations);
        cudaEventRecord(stop);
        if (cudaEventSynchronize(stop) != cudaSuccess) { best = -1.0; break; }
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, start, stop);
        const double flops = 2.0 * 8.0 * iterations * (double)blocks * threads;
        const double tflops = flops / (ms * 1.0e-3) * 1.0e-12;
        if (rep > 0 && tflops > best) best = tflops;
    }
    cudaEventDestroy(start);
    cudaEventDestroy(stop);
    cudaFree(out);
    return best;
}
