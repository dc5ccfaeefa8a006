// Micro probe: what the L2 charges for scattered fp32 reductions (red.global.add) on B200.
// Decides how the d loss / d volume scatter of the MLP-training step (BASELINE config 4) should be arranged:
// is the cost per lane, per 32-byte sector, or per instruction, and do the sm_90+ vector forms
// (red.global.add.v2.f32 / .v4.f32) buy anything?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o red_probe red_probe.cu && ./red_probe
//
// Every pattern issues the same number of warp instructions over a 64 MiB buffer (the size of a 256^3 gradient
// volume, L2-resident); addresses come from a per-thread LCG so that no two instructions of a warp repeat a sector.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

constexpr int ITERS = 64;

__device__ __forceinline__ uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s; }
__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

__device__ __forceinline__ void red1(float* p, float v) { asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }
__device__ __forceinline__ void red2(float* p, float v) { asm volatile("red.global.add.v2.f32 [%0], {%1, %1};" ::"l"(p), "f"(v) : "memory"); }
__device__ __forceinline__ void red4(float* p, float v) { asm volatile("red.global.add.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(p), "f"(v) : "memory"); }

// mode: 0 scalar, every lane its own random sector                         (32 sectors / instruction, 1 float each)
//       1 scalar, 8 lanes fill one random sector                           ( 4 sectors / instruction, 8 floats each)
//       2 scalar, 32 lanes fill one random 128-byte line                   ( 4 sectors, one line)
//       3 scalar, lane pairs hit the SAME random address                   (16 sectors, 2-way address conflict)
//       4 v2, every lane its own random sector                             (32 sectors, 2 floats each)
//       5 v4, every lane its own random sector                             (32 sectors, 4 floats each)
//       6 v4, lane pairs fill one random sector                            (16 sectors, 8 floats each)
//       7 v4, 8 lanes fill one random 128-byte line                        ( 4 sectors per line, 4 lines / instruction)
//       8 scalar, 4 lanes share one random sector (distinct floats)        ( 8 sectors, 4 floats each)
//       9 scalar, 2 lanes share one random sector (distinct floats)        (16 sectors, 2 floats each)
__global__ void probe(float* buf, uint32_t n_sectors, int mode, uint32_t seed) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31, warp = tid >> 5;
    uint32_t s = mix(seed ^ (tid * 2654435761u));
    uint32_t ws = mix(seed ^ (warp * 40503u + 977u));
    for (int it = 0; it < ITERS; ++it) {
        uint32_t r = lcg(s), w = lcg(ws);
        uint32_t sector;
        float* p;
        switch (mode) {
            case 0: sector = r % n_sectors; p = buf + (size_t)sector * 8 + (r >> 29); red1(p, 1.f); break;
            case 1: sector = mix(w + (lane >> 3)) % n_sectors; p = buf + (size_t)sector * 8 + (lane & 7); red1(p, 1.f); break;
            case 2: sector = (mix(w) % (n_sectors / 4)) * 4; p = buf + (size_t)sector * 8 + lane; red1(p, 1.f); break;
            case 3: sector = mix(w + (lane >> 1)) % n_sectors; p = buf + (size_t)sector * 8 + ((w >> 7) & 7); red1(p, 1.f); break;
            case 4: sector = r % n_sectors; p = buf + (size_t)sector * 8 + ((r >> 30) << 1); red2(p, 1.f); break;
            case 5: sector = r % n_sectors; p = buf + (size_t)sector * 8 + ((r >> 31) << 2); red4(p, 1.f); break;
            case 6: sector = mix(w + (lane >> 1)) % n_sectors; p = buf + (size_t)sector * 8 + ((lane & 1) << 2); red4(p, 1.f); break;
            case 7: sector = (mix(w + (lane >> 3)) % (n_sectors / 4)) * 4; p = buf + (size_t)sector * 8 + ((lane & 7) << 2); red4(p, 1.f); break;
            case 8: sector = mix(w + (lane >> 2)) % n_sectors; p = buf + (size_t)sector * 8 + (lane & 3) * 2; red1(p, 1.f); break;
            default: sector = mix(w + (lane >> 1)) % n_sectors; p = buf + (size_t)sector * 8 + (lane & 1) * 4; red1(p, 1.f); break;
        }
    }
}

// the read-side reference: random-sector loads with the same address stream as mode 0
__global__ void probe_read(const float* buf, uint32_t n_sectors, uint32_t seed, float* sink) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t s = mix(seed ^ (tid * 2654435761u));
    float acc = 0.f;
#pragma unroll 8
    for (int it = 0; it < ITERS; ++it) {
        uint32_t r = lcg(s);
        acc += __ldg(buf + (size_t)(r % n_sectors) * 8 + (r >> 29));
    }
    sink[tid] = acc;
}

int main() {
    const size_t bytes = 64u << 20;
    const uint32_t n_sectors = bytes / 32;
    float *buf, *sink;
    cudaMalloc(&buf, bytes);
    cudaMemset(buf, 0, bytes);
    const int blocks = 148 * 16, threads = 256;
    cudaMalloc(&sink, (size_t)blocks * threads * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const char* names[] = {"scalar, 32 random sectors/instr (1 float each)", "scalar, 4 random sectors/instr (8 floats each)",
                           "scalar, 1 random 128-B line/instr", "scalar, 16 sectors/instr, pairs on the SAME address",
                           "v2, 32 random sectors/instr", "v4, 32 random sectors/instr", "v4, 16 random sectors/instr (full sectors)",
                           "v4, 4 random lines/instr (full lines)", "scalar, 8 random sectors/instr (4 floats each)",
                           "scalar, 16 random sectors/instr (2 floats each)"};
    const int sectors_per_instr[] = {32, 4, 4, 16, 32, 32, 16, 16, 8, 16};
    const int floats_per_lane[] = {1, 1, 1, 1, 2, 4, 4, 4, 1, 1};
    const double instrs = (double)blocks * threads / 32 * ITERS;
    printf("| pattern | ms | G warp-instr/s | G lane-ops/s | G sectors/s | G floats/s |\n|---|---:|---:|---:|---:|---:|\n");
    for (int mode = 0; mode < 10; ++mode) {
        for (int w = 0; w < 2; ++w) probe<<<blocks, threads>>>(buf, n_sectors, mode, 17 + w);
        cudaEventRecord(e0);
        const int reps = 5;
        for (int r = 0; r < reps; ++r) probe<<<blocks, threads>>>(buf, n_sectors, mode, 100 + r);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        ms /= reps;
        double t = ms * 1e-3;
        printf("| %s | %.4f | %.2f | %.1f | %.1f | %.1f |\n", names[mode], ms, instrs / t / 1e9, instrs * 32 / t / 1e9,
               instrs * sectors_per_instr[mode] / t / 1e9, instrs * 32 * floats_per_lane[mode] / t / 1e9);
    }
    for (int w = 0; w < 2; ++w) probe_read<<<blocks, threads>>>(buf, n_sectors, 17 + w, sink);
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) probe_read<<<blocks, threads>>>(buf, n_sectors, 100 + r, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= 5;
    printf("| ld.global.nc, 32 random sectors/instr (reference) | %.4f | %.2f | %.1f | %.1f | %.1f |\n", ms, instrs / (ms * 1e-3) / 1e9,
           instrs * 32 / (ms * 1e-3) / 1e9, instrs * 32 / (ms * 1e-3) / 1e9, instrs * 32 / (ms * 1e-3) / 1e9);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
