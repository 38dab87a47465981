// fp64_peak.cu -- measures the FP64 pipes of the device this runs on: DMMA.8x8x4 (mma.sync f64)
// and DFMA issue rates from registers.  Used once per pool to calibrate the FP64 roofline
// denominator (MEASURED_PEAKS.json only carries HBM and bf16 numbers).
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dmma_loop(double* out, int iters) {
    double c[16][2];
    for (int i = 0; i < 16; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
    for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void dfma_loop(double* out, int iters) {
    double c[16];
    for (int i = 0; i < 16; ++i) c[i] = i;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0;
    for (int i = 0; i < 16; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    double* out; cudaMalloc(&out, sizeof(double) * sms * 4 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int warps = 4; warps <= 16; warps *= 2) {
        const int iters = 20000;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            dmma_loop<<<sms, warps * 32>>>(out, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double flops = 2.0 * 256 * 16.0 * iters * warps * sms;
            if (rep) printf("DMMA.8x8x4  warps/SM=%2d  %.2f TFLOP/s  (%.3f ms)\n", warps, flops / ms * 1e-9, ms);
        }
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            dfma_loop<<<sms, warps * 32>>>(out, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double flops = 2.0 * 32 * 16.0 * iters * warps * sms;
            if (rep) printf("DFMA        warps/SM=%2d  %.2f TFLOP/s  (%.3f ms)\n", warps, flops / ms * 1e-9, ms);
        }
    }
    printf("device %s, %d SMs, clock %d kHz\n", p.name, sms, p.clockRate);
    return 0;
}
