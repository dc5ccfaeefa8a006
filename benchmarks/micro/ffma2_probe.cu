// Microbenchmark: does the packed FP32 FMA of sm_100 (FFMA2, `fma.rn.f32x2`) free issue slots?
// Four kernels with the same number of FP32 FMAs per thread:
//   scalar      : FFMA only                       packed      : FFMA2 only
//   scalar_int  : FFMA + as many integer ops      packed_int  : FFMA2 + the same integer ops
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o benchmarks/micro/ffma2_probe benchmarks/micro/ffma2_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096, CHAINS = 8;

template <bool PACKED, bool WITH_INT>
__global__ void __launch_bounds__(256) probe(float* out, unsigned* iout, float a, float b, unsigned salt) {
    float x[CHAINS];
    unsigned u[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) { x[c] = threadIdx.x * 1e-3f + c; u[c] = threadIdx.x * 2654435761u + c + salt; }
    for (int it = 0; it < ITERS; ++it) {
        if (PACKED) {
#pragma unroll
            for (int c = 0; c < CHAINS; c += 2) {
                float2 v = __ffma2_rn(make_float2(x[c], x[c + 1]), make_float2(a, a), make_float2(b, b));
                x[c] = v.x; x[c + 1] = v.y;
            }
        } else {
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) x[c] = __fmaf_rn(x[c], a, b);
        }
        if (WITH_INT) {
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) u[c] = (u[c] ^ (u[c] >> 7)) + salt;     // LOP3/SHF + IADD
        }
    }
    float s = 0.f;
    unsigned t = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) { s += x[c]; t += u[c]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    iout[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

template <bool P, bool I>
static float run(const char* name, float* out, unsigned* iout) {
    const int blocks = 148 * 8, threads = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 3; ++w) probe<P, I><<<blocks, threads>>>(out, iout, 0.999f, 0.001f, 17u);
    cudaEventRecord(e0);
    for (int r = 0; r < 10; ++r) probe<P, I><<<blocks, threads>>>(out, iout, 0.999f, 0.001f, 17u);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= 10.f;
    double fma = (double)blocks * threads * ITERS * CHAINS;
    printf("{\"kernel\": \"%s\", \"ms\": %.4f, \"fp32_fma_per_s\": %.4e, \"fma_per_clk_per_sm_at_1965MHz\": %.1f}\n", name, ms,
           fma / (ms * 1e-3), fma / (ms * 1e-3) / 148.0 / 1.965e9);
    return ms;
}

int main() {
    float* out; unsigned* iout;
    cudaMalloc(&out, 148 * 8 * 256 * 4); cudaMalloc(&iout, 148 * 8 * 256 * 4);
    run<false, false>("scalar", out, iout);
    run<true, false>("packed", out, iout);
    run<false, true>("scalar_int", out, iout);
    run<true, true>("packed_int", out, iout);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
