// Microbenchmark: issue rate of DFMA / DADD / DMUL on sm_100a (register-only loops, 8 independent chains per thread).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_rate fp64_rate.cu && ./fp64_rate
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(double* out, double a, double b, int iters) {
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = a + threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) x[i] = fma(x[i], a, b);
                if (MODE == 1) x[i] = x[i] + b;
                if (MODE == 2) x[i] = x[i] * a;
                if (MODE == 3) { x[i] = (r & 1) ? x[i] + b : fma(x[i], a, b); }        // 50 % DADD, 50 % DFMA
                if (MODE == 4) { float f = (float)x[i]; x[i] = (double)f + b; }       // F2F pair + DADD
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    if (s == 12345.678) out[0] = s;
}

struct WW { double w[16]; };
// MODE 0: weights as uniform-register / constant operands (what the convolution kernels do); 1: weights forced into registers
template <int MODE>
__global__ void __launch_bounds__(256) kw(double* out, WW w, int iters) {
    double x[8], wr[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = 1.0 + threadIdx.x * 1e-9 + i;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        wr[i] = w.w[i];
        if (MODE == 1) asm volatile("" : "+d"(wr[i]));
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = fma(x[i], wr[(4 * r + i) & 15], wr[(r + 2 * i + 1) & 15]);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    if (s == 12345.678) out[0] = s;
}

template <int MODE>
void runw(const char* name, double* d) {
    int dev_sms; cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, 0);
    const int iters = 4096;
    WW w; for (int i = 0; i < 16; ++i) w.w[i] = 1.0 + 1e-7 * i;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int blocks_per_sm : {1, 2, 4}) {
        kw<MODE><<<dev_sms * blocks_per_sm, 256>>>(d, w, 16);
        cudaEventRecord(e0);
        kw<MODE><<<dev_sms * blocks_per_sm, 256>>>(d, w, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double lane_ops = (double)dev_sms * blocks_per_sm * 256 * iters * 32.0;
        int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
        printf("%-22s warps/SM %2d  %.3f ms  = %.1f lanes/clk/SM\n", name, blocks_per_sm * 8, ms, lane_ops / (ms * 1e-3) / dev_sms / (clk * 1e3));
    }
}

template <int MODE>
void run(const char* name, int ops_per_inner, double* d) {
    int dev_sms; cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, 0);
    const int iters = 4096;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int blocks_per_sm : {1, 2, 4, 8}) {
        k<MODE><<<dev_sms * blocks_per_sm, 256>>>(d, 1.0000001, 1e-9, 16);
        cudaEventRecord(e0);
        k<MODE><<<dev_sms * blocks_per_sm, 256>>>(d, 1.0000001, 1e-9, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double lane_ops = (double)dev_sms * blocks_per_sm * 256 * iters * 32.0 * ops_per_inner;
        int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
        printf("%-22s warps/SM %2d  %.3f ms  %.2f T lane-ops/s  = %.1f lanes/clk/SM at %d MHz\n", name, blocks_per_sm * 8, ms,
               lane_ops / ms / 1e9, lane_ops / (ms * 1e-3) / dev_sms / (clk * 1e3), clk / 1000);
    }
}

int main() {
    double* d; cudaMalloc(&d, 8);
    run<0>("DFMA", 1, d);
    run<1>("DADD", 1, d);
    run<2>("DMUL", 1, d);
    run<3>("DADD+DFMA 50/50", 1, d);
    run<4>("F2F+F2F+DADD (per 3)", 1, d);
    runw<0>("DFMA uniform weights", d);
    runw<1>("DFMA register weights", d);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
