// Impedance MLP forward with layer 2 on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a only.
//
// Of Linear(1,32)-ReLU-Linear(32,32)-ReLU-Linear(32,1) only layer 2 is a dense contraction:
//     H2pre[128 voxels x 32] = H1[128 x 32] * W2^T[32 x 32]          per 128-voxel tile.
// A single TF32 pass cannot hold the renderer's tolerance (impedances enter the reflection coefficient as
// a difference of nearly equal numbers), so the product is evaluated as the 3xTF32 split
//     H1 W2^T ~= H1_hi W2_hi^T + H1_lo W2_hi^T + H1_hi W2_lo^T ,   x_hi = x & 0xffffe000, x_lo = x - x_hi,
// accumulated in fp32 in TMEM (relative error ~2^-21 per product).  Layers 1 and 3 (an outer and an inner
// product) stay on the CUDA cores: each thread owns one voxel, builds its row of H1 (hi and lo) directly in
// the canonical K-major shared-memory layout the tensor core reads (8-row x 16-byte core matrices, no
// swizzle), and after the MMAs reads its row of H2pre back from TMEM with one tcgen05.ld (32x32b.x32).
//
// One elected thread issues 12 tcgen05.mma (M=128, N=32, K=8) per tile and commits them to an mbarrier.
#include "common.cuh"
#include "launch.h"

namespace diffus {

namespace {

constexpr int TILE_M = 128, HID = 32;
constexpr int OFF_W1 = 0, OFF_B1 = 32, OFF_W2 = 64, OFF_B2 = 64 + 1024, OFF_W3 = OFF_B2 + 32, OFF_B3 = OFF_W3 + 32;
// canonical K-major layout, no swizzle: [k_chunk of 4 floats][row group of 8][row in group][4 floats]
constexpr uint32_t A_SBO = 128, A_LBO = (TILE_M / 8) * 128;     // bytes: next 8-row group, next 16-byte K chunk
constexpr uint32_t B_SBO = 128, B_LBO = (HID / 8) * 128;
constexpr int A_FLOATS = TILE_M * HID, B_FLOATS = HID * HID;
constexpr int SMEM_FLOATS = 2 * A_FLOATS + 2 * B_FLOATS + 4 * HID + 4;
constexpr uint32_t TMEM_COLS = 32;
// instruction descriptor: D=F32 (1<<4), A=TF32 (2<<7), B=TF32 (2<<10), both K-major, N>>3 at [17,23), M>>4 at [24,29)
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(HID >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = (uint64_t)((addr >> 4) & 0x3fffu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
    d |= (uint64_t)1 << 46;                                     // descriptor version for sm_100
    return d;                                                   // base_offset 0, layout_type 0 = no swizzle
}

__device__ __forceinline__ float tf32_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xffffe000u); }

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(a_desc), "l"(b_desc), "r"(IDESC), "r"(accumulate)
        : "memory");
}

}  // namespace

__global__ void __launch_bounds__(TILE_M) mlp_fwd_tc_kernel(const float* __restrict__ params, const float* __restrict__ x,
                                                            const uint8_t* __restrict__ mask, int64_t n, float out_scale,
                                                            float fill, float* __restrict__ out) {
    extern __shared__ __align__(128) float smem[];
    float* a_hi = smem;
    float* a_lo = a_hi + A_FLOATS;
    float* b_hi = a_lo + A_FLOATS;
    float* b_lo = b_hi + B_FLOATS;
    float* w1 = b_lo + B_FLOATS;
    float* b1 = w1 + HID;
    float* b2 = b1 + HID;
    float* w3 = b2 + HID;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_slot;
    const int tid = threadIdx.x, warp = tid >> 5;

    if (tid < HID) {
        w1[tid] = params[OFF_W1 + tid];
        b1[tid] = params[OFF_B1 + tid];
        b2[tid] = params[OFF_B2 + tid];
        w3[tid] = params[OFF_W3 + tid];
    }
    const float b3 = params[OFF_B3];
    // W2 (out j, in i) is already "N x K, K-major": B[n][k] = W2[n][k]
    for (int e = tid; e < B_FLOATS; e += TILE_M) {
        int nn = e >> 5, k = e & 31;
        float v = params[OFF_W2 + e], hi = tf32_hi(v);
        int idx = (k >> 2) * (B_LBO / 4) + (nn >> 3) * (B_SBO / 4) + (nn & 7) * 4 + (k & 3);
        b_hi[idx] = hi;
        b_lo[idx] = v - hi;
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_slot)), "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;\n");
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");      // W2 tiles: generic-proxy writes -> tensor-core reads
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n");
    const uint32_t tmem = tmem_base_slot;
    const uint32_t a_hi_s = smem_u32(a_hi), a_lo_s = smem_u32(a_lo), b_hi_s = smem_u32(b_hi), b_lo_s = smem_u32(b_lo);
    const uint32_t bar_s = smem_u32(&bar);
    uint32_t phase = 0;

    const int64_t n_tiles = (n + TILE_M - 1) / TILE_M;
    // the next tile's input is fetched one iteration ahead: its DRAM latency hides behind this tile's work
    float x_next = 0.f;
    {
        const int64_t i0 = (int64_t)blockIdx.x * TILE_M + tid;
        if (blockIdx.x < n_tiles && i0 < n) x_next = __ldg(x + i0);
    }
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t idx = tile * TILE_M + tid;
        const float xv = x_next;
        {
            const int64_t inext = (tile + gridDim.x) * TILE_M + tid;
            x_next = (tile + gridDim.x < n_tiles && inext < n) ? __ldg(x + inext) : 0.f;
        }
        // layer 1 on the CUDA cores, written straight into the tensor core's operand layout
        float* row_hi = a_hi + (tid >> 3) * (A_SBO / 4) + (tid & 7) * 4;
        float* row_lo = a_lo + (tid >> 3) * (A_SBO / 4) + (tid & 7) * 4;
#pragma unroll
        for (int kc = 0; kc < HID / 4; ++kc) {
            const float4 w4 = ((const float4*)w1)[kc], c4 = ((const float4*)b1)[kc];     // broadcast 16-byte reads
            const float wv[4] = {w4.x, w4.y, w4.z, w4.w}, cv[4] = {c4.x, c4.y, c4.z, c4.w};
            float h[4], hi[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                h[j] = fmaxf(fmaf(wv[j], xv, cv[j]), 0.f);
                hi[j] = tf32_hi(h[j]);
            }
            *(float4*)(row_hi + kc * (A_LBO / 4)) = make_float4(hi[0], hi[1], hi[2], hi[3]);
            *(float4*)(row_lo + kc * (A_LBO / 4)) = make_float4(h[0] - hi[0], h[1] - hi[1], h[2] - hi[2], h[3] - hi[3]);
        }
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;\n");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;\n");
            uint32_t acc = 0;
#pragma unroll
            for (int pass = 0; pass < 3; ++pass) {
                const uint32_t a_s = pass == 1 ? a_lo_s : a_hi_s, b_s = pass == 2 ? b_lo_s : b_hi_s;
#pragma unroll
                for (int j = 0; j < HID / 8; ++j) {                    // K = 8 per MMA = two 16-byte chunks
                    mma_tf32(tmem, smem_desc(a_s + j * 2 * A_LBO, A_LBO, A_SBO), smem_desc(b_s + j * 2 * B_LBO, B_LBO, B_SBO), acc);
                    acc = 1;
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar_s) : "memory");
        }
        // wait for the accumulator (bounded: a mis-programmed pipeline must fail, not hang the GPU)
        uint32_t done = 0;
        for (int spin = 0; spin < (1 << 24) && !done; ++spin)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                         : "=r"(done) : "r"(bar_s), "r"(phase) : "memory");
        if (!done) __trap();
        phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;\n");
        uint32_t v[32];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);     // this warp's 32 TMEM lanes, 32 columns
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
              "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
              "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        // layer 3 on the CUDA cores
        float acc3a = b3, acc3b = 0.f;
#pragma unroll
        for (int j4 = 0; j4 < HID / 4; ++j4) {
            const float4 c4 = ((const float4*)b2)[j4], w4 = ((const float4*)w3)[j4];
            acc3a = fmaf(w4.x, fmaxf(__uint_as_float(v[4 * j4]) + c4.x, 0.f), acc3a);
            acc3b = fmaf(w4.y, fmaxf(__uint_as_float(v[4 * j4 + 1]) + c4.y, 0.f), acc3b);
            acc3a = fmaf(w4.z, fmaxf(__uint_as_float(v[4 * j4 + 2]) + c4.z, 0.f), acc3a);
            acc3b = fmaf(w4.w, fmaxf(__uint_as_float(v[4 * j4 + 3]) + c4.w, 0.f), acc3b);
        }
        const float acc3 = acc3a + acc3b;
        if (idx < n) out[idx] = (mask && !mask[idx]) ? fill : out_scale * acc3;
        // the next tile overwrites the operand tiles and the accumulator
        asm volatile("tcgen05.fence::before_thread_sync;\n");
        __syncthreads();
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(TMEM_COLS));
}

cudaError_t launch_mlp_fwd_tc(const float* params, const float* x, const uint8_t* mask, int64_t n, float out_scale,
                              float fill, float* out, cudaStream_t st) {
    size_t smem = SMEM_FLOATS * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(mlp_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int64_t tiles = (n + TILE_M - 1) / TILE_M;
    unsigned grid = (unsigned)max((int64_t)1, min(tiles, (int64_t)148 * 5));     // 5 CTAs per SM fit (43 KB of shared memory each)
    mlp_fwd_tc_kernel<<<grid, TILE_M, smem, st>>>(params, x, mask, n, out_scale, fill, out);
    return cudaGetLastError();
}


// =========================================================================================
// Backward on the tensor cores.
//
// Per 128-voxel tile (thread = voxel), with H1 = relu(W1 x + b1) and g = d loss / d out:
//   G1  S    = H1  * W2^T          recompute of layer 2's pre-activation      tcgen05, K = 32 units
//   G2  DH1p = DH2 * W2            DH2 = g w3^T [S + b2 > 0]                   tcgen05, K = 32 units
//   G3  dW2 += DH2^T * H1          a contraction over VOXELS                   mma.sync m16n8k8, K = 32 voxels per warp
// all three as 3xTF32 splits (hi*hi + hi*lo + lo*hi).  G1 / G2 read the operand tiles through shared-memory
// descriptors with the accumulators in TMEM, like the forward.  G3 cannot: its K dimension is the voxel index, the
// tiles are unit-contiguous, and tcgen05's kind::tf32 does not read transposed (MN-major) operands -- the accumulator
// simply stays 0 (benchmarks/micro/umma_mn_probe.cu, profiles/r1_micro_probes.md).  The warp-level mma.sync takes its
// fragments from registers, so each warp gathers them from the same tiles with plain shared-memory loads (the group
// stride of the tiles is padded to 16 (mod 32) words, which makes those loads conflict free) and keeps its 32x32
// partial of dW2 in 32 registers across all its tiles; the tensor core works on G2 of the same tile meanwhile.
// The vector gradients (dW3, db3, db2, db1, dW1) are reduced over the 32 voxels of a warp with a 31-shuffle
// transpose-reduction and kept in one register per lane.
// =========================================================================================
namespace {

constexpr int GS = TILE_M * 4 + 16;                       // floats between groups of 4 units: [group][voxel][4], 16 (mod 32)
constexpr uint32_t GROUP_BYTES = GS * 4;
constexpr int BT_TILE_FLOATS = 16 * GS;                   // hi groups 0-7, lo groups 8-15
constexpr uint32_t TMEM_COLS_BWD = 64;                    // S [0,32) | DH1 [32,64)
constexpr uint32_t TM_S = 0, TM_DH1 = 32;
constexpr int BT_SMEM_FLOATS = 2 * BT_TILE_FLOATS + 4 * B_FLOATS + 4 * HID + 8;

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t v[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar_s, uint32_t phase) {
    uint32_t done = 0;
    for (int spin = 0; spin < (1 << 24) && !done; ++spin)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(bar_s), "r"(phase) : "memory");
    if (!done) __trap();                    // a mis-programmed pipeline must fail loudly, not hang the GPU
}

// D (16x8, fp32) += A (16x8, row) * B (8x8, col), TF32 operands
__device__ __forceinline__ void mma_16x8x8_tf32(float d[4], const uint32_t a[4], const uint32_t b[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// sum over the 32 lanes of val[j] for every j; lane l returns the total of index l (31 shuffles)
__device__ __forceinline__ float transpose_reduce(float val[32], int lane) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            float keep = upper ? val[i + off] : val[i];
            float send = upper ? val[i] : val[i + off];
            val[i] = keep + __shfl_xor_sync(FULL, send, off);
        }
    }
    return val[0];
}

}  // namespace

__global__ void __launch_bounds__(TILE_M) mlp_bwd_tc_kernel(const float* __restrict__ params, const float* __restrict__ x,
                                                            const uint8_t* __restrict__ mask, const float* __restrict__ grad_out,
                                                            int64_t n, float out_scale, float* __restrict__ block_partials) {
    extern __shared__ __align__(128) float smem[];
    float* dh2_t = smem;                              // [16 groups][GS]: DH2 hi (groups 0-7) and lo (8-15)
    float* h1_t = dh2_t + BT_TILE_FLOATS;             // same for H1
    float* w2_hi = h1_t + BT_TILE_FLOATS;             // B of G1: [n = j][k = i]
    float* w2_lo = w2_hi + B_FLOATS;
    float* w2t_hi = w2_lo + B_FLOATS;                 // B of G2: [n = i][k = j]
    float* w2t_lo = w2t_hi + B_FLOATS;
    float* w1 = w2t_lo + B_FLOATS;
    float* b1 = w1 + HID;
    float* b2 = b1 + HID;
    float* w3 = b2 + HID;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_slot;
    __shared__ float red[DIFFUS_MLP_NPARAMS];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int grp = lane >> 2, tig = lane & 3;        // mma.sync fragment coordinates

    if (tid < HID) {
        w1[tid] = params[OFF_W1 + tid];
        b1[tid] = params[OFF_B1 + tid];
        b2[tid] = params[OFF_B2 + tid];
        w3[tid] = params[OFF_W3 + tid];
    }
    for (int e = tid; e < B_FLOATS; e += TILE_M) {
        int j = e >> 5, i = e & 31;                   // W2[j][i]
        float v = params[OFF_W2 + e], hi = tf32_hi(v);
        int idx = (i >> 2) * (B_LBO / 4) + (j >> 3) * (B_SBO / 4) + (j & 7) * 4 + (i & 3);     // rows j, K = i
        int idt = (j >> 2) * (B_LBO / 4) + (i >> 3) * (B_SBO / 4) + (i & 7) * 4 + (j & 3);     // rows i, K = j
        w2_hi[idx] = hi;
        w2_lo[idx] = v - hi;
        w2t_hi[idt] = hi;
        w2t_lo[idt] = v - hi;
    }
    for (int i = tid; i < DIFFUS_MLP_NPARAMS; i += TILE_M) red[i] = 0.f;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_slot)), "n"(TMEM_COLS_BWD));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;\n");
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n");
    const uint32_t tmem = tmem_base_slot;
    const uint32_t dh2_s = smem_u32(dh2_t), h1_s = smem_u32(h1_t);
    const uint32_t w2_hi_s = smem_u32(w2_hi), w2_lo_s = smem_u32(w2_lo), w2t_hi_s = smem_u32(w2t_hi), w2t_lo_s = smem_u32(w2t_lo);
    const uint32_t bar_s = smem_u32(&bar);
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    uint32_t phase = 0;
    float acc_w3 = 0.f, acc_b2 = 0.f, acc_b1 = 0.f, acc_w1 = 0.f, acc_b3 = 0.f;      // lane j / i of each warp
    // this warp's 32x32 partial of dW2: the mma.sync accumulator, folded every FLUSH tiles into a second set of registers
    // with round-to-nearest adds (the tensor core's own fp32 accumulation may truncate; a few hundred terms per segment
    // keep that below 1e-5 whatever it does)
    constexpr int FLUSH = 8;
    float dw2[2][4][4], dw2_sum[2][4][4];
    int since_flush = 0;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int c = 0; c < 4; ++c) dw2[mt][nt][c] = dw2_sum[mt][nt][c] = 0.f;

    const int64_t n_tiles = (n + TILE_M - 1) / TILE_M;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t idx = tile * TILE_M + tid;
        float xv = 0.f, g = 0.f;
        if (idx < n) {
            xv = __ldg(x + idx);
            g = __ldg(grad_out + idx) * out_scale;
            if (mask && !mask[idx]) g = 0.f;
        }
        if (!__syncthreads_or(g != 0.f)) continue;                        // voxels no ray touched: nothing to add
        // ---- layer 1 -> H1 tiles (operand layout: [group of 4 units][voxel][4]) ----
#pragma unroll
        for (int kc = 0; kc < HID / 4; ++kc) {
            const float4 w4 = ((const float4*)w1)[kc], c4 = ((const float4*)b1)[kc];
            const float wv[4] = {w4.x, w4.y, w4.z, w4.w}, cv[4] = {c4.x, c4.y, c4.z, c4.w};
            float h[4], hi[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                h[j] = fmaxf(fmaf(wv[j], xv, cv[j]), 0.f);
                hi[j] = tf32_hi(h[j]);
            }
            *(float4*)(h1_t + kc * GS + tid * 4) = make_float4(hi[0], hi[1], hi[2], hi[3]);
            *(float4*)(h1_t + (8 + kc) * GS + tid * 4) = make_float4(h[0] - hi[0], h[1] - hi[1], h[2] - hi[2], h[3] - hi[3]);
        }
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;\n");
        __syncthreads();
        // ---- G1: S = H1 W2^T ----
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;\n");
            uint32_t acc = 0;
#pragma unroll
            for (int pass = 0; pass < 3; ++pass) {
                const uint32_t a_s = h1_s + (pass == 1 ? 8 * GROUP_BYTES : 0), b_s = pass == 2 ? w2_lo_s : w2_hi_s;
#pragma unroll
                for (int j = 0; j < HID / 8; ++j) {
                    mma_tf32(tmem + TM_S, smem_desc(a_s + j * 2 * GROUP_BYTES, GROUP_BYTES, 128),
                             smem_desc(b_s + j * 2 * B_LBO, B_LBO, B_SBO), acc);
                    acc = 1;
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar_s) : "memory");
        }
        mbar_wait(bar_s, phase);
        phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;\n");
        // ---- epilogue 1: h2, DH2 = g w3 [s > 0] -> DH2 tiles; dW3, db3, db2 ----
        {
            uint32_t sv[32];
            tmem_ld32(tmem + lane_base + TM_S, sv);
            float p3[32];
#pragma unroll
            for (int jc = 0; jc < HID / 4; ++jc) {
                const float4 c4 = ((const float4*)b2)[jc], w4 = ((const float4*)w3)[jc];
                const float cv[4] = {c4.x, c4.y, c4.z, c4.w}, wv[4] = {w4.x, w4.y, w4.z, w4.w};
                float d[4], dhi[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float sj = __uint_as_float(sv[4 * jc + j]) + cv[j];
                    p3[4 * jc + j] = g * fmaxf(sj, 0.f);
                    d[j] = sj > 0.f ? g * wv[j] : 0.f;
                    dhi[j] = tf32_hi(d[j]);
                    sv[4 * jc + j] = __float_as_uint(d[j]);              // keep DH2 for db2
                }
                *(float4*)(dh2_t + jc * GS + tid * 4) = make_float4(dhi[0], dhi[1], dhi[2], dhi[3]);
                *(float4*)(dh2_t + (8 + jc) * GS + tid * 4) = make_float4(d[0] - dhi[0], d[1] - dhi[1], d[2] - dhi[2], d[3] - dhi[3]);
            }
            acc_w3 += transpose_reduce(p3, lane);
#pragma unroll
            for (int j = 0; j < HID; ++j) p3[j] = __uint_as_float(sv[j]);
            acc_b2 += transpose_reduce(p3, lane);
            acc_b3 += warp_sum(g);
        }
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;\n");
        __syncthreads();
        // ---- G2 on the tensor core: DH1pre = DH2 W2 ----
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;\n");
            uint32_t acc = 0;
#pragma unroll
            for (int pass = 0; pass < 3; ++pass) {
                const uint32_t a_s = dh2_s + (pass == 1 ? 8 * GROUP_BYTES : 0), b_s = pass == 2 ? w2t_lo_s : w2t_hi_s;
#pragma unroll
                for (int j = 0; j < HID / 8; ++j) {
                    mma_tf32(tmem + TM_DH1, smem_desc(a_s + j * 2 * GROUP_BYTES, GROUP_BYTES, 128),
                             smem_desc(b_s + j * 2 * B_LBO, B_LBO, B_SBO), acc);
                    acc = 1;
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar_s) : "memory");
        }
        // ---- G3 meanwhile, on this warp's 32 voxels: dW2[m][n] += sum_v DH2[v][m] H1[v][n] ----
        // fragment element (unit u, voxel v) of either tile sits at (u >> 2) * GS + v * 4 + (u & 3)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            const int v0 = (warp * 32 + ks * 8 + tig) * 4;             // k-slot tig; slot tig + 4 is 16 floats on
            uint32_t a_hi[2][4], a_lo[2][4], b_hi[4][2], b_lo[4][2];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                const int m0 = 16 * mt + grp;                          // rows m0 and m0 + 8 = two groups on
                const float* ph = dh2_t + (m0 >> 2) * GS + v0 + (m0 & 3);
                a_hi[mt][0] = __float_as_uint(ph[0]);          a_hi[mt][1] = __float_as_uint(ph[2 * GS]);
                a_hi[mt][2] = __float_as_uint(ph[16]);         a_hi[mt][3] = __float_as_uint(ph[2 * GS + 16]);
                const float* pl = ph + 8 * GS;
                a_lo[mt][0] = __float_as_uint(pl[0]);          a_lo[mt][1] = __float_as_uint(pl[2 * GS]);
                a_lo[mt][2] = __float_as_uint(pl[16]);         a_lo[mt][3] = __float_as_uint(pl[2 * GS + 16]);
            }
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const int n0 = 8 * nt + grp;
                const float* ph = h1_t + (n0 >> 2) * GS + v0 + (n0 & 3);
                b_hi[nt][0] = __float_as_uint(ph[0]);          b_hi[nt][1] = __float_as_uint(ph[16]);
                b_lo[nt][0] = __float_as_uint(ph[8 * GS]);     b_lo[nt][1] = __float_as_uint(ph[8 * GS + 16]);
            }
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    mma_16x8x8_tf32(dw2[mt][nt], a_lo[mt], b_hi[nt]);  // small terms first
                    mma_16x8x8_tf32(dw2[mt][nt], a_hi[mt], b_lo[nt]);
                    mma_16x8x8_tf32(dw2[mt][nt], a_hi[mt], b_hi[nt]);
                }
        }
        if (++since_flush == FLUSH) {
            since_flush = 0;
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                    for (int c = 0; c < 4; ++c) { dw2_sum[mt][nt][c] += dw2[mt][nt][c]; dw2[mt][nt][c] = 0.f; }
        }
        mbar_wait(bar_s, phase);
        phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;\n");
        // ---- epilogue 2: dh1 = DH1pre [h1 > 0]; db1, dW1 ----
        {
            uint32_t dv[32];
            tmem_ld32(tmem + lane_base + TM_DH1, dv);
            float q1[32], qx[32];
#pragma unroll
            for (int kc = 0; kc < HID / 4; ++kc) {
                const float4 w4 = ((const float4*)w1)[kc], c4 = ((const float4*)b1)[kc];
                const float wv[4] = {w4.x, w4.y, w4.z, w4.w}, cv[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float dh1 = fmaf(wv[j], xv, cv[j]) > 0.f ? __uint_as_float(dv[4 * kc + j]) : 0.f;
                    q1[4 * kc + j] = dh1;
                    qx[4 * kc + j] = dh1 * xv;
                }
            }
            acc_b1 += transpose_reduce(q1, lane);
            acc_w1 += transpose_reduce(qx, lane);
        }
        asm volatile("tcgen05.fence::before_thread_sync;\n");
        __syncthreads();                                   // operand tiles and the S / DH1 accumulators are reused
    }
    // ---- block partials: the four warps add their vectors and their dW2 fragments in a fixed order ----
    __syncthreads();
    for (int wv = 0; wv < TILE_M / 32; ++wv) {
        if (warp == wv) {
            red[OFF_W3 + lane] += acc_w3;
            red[OFF_B2 + lane] += acc_b2;
            red[OFF_B1 + lane] += acc_b1;
            red[OFF_W1 + lane] += acc_w1;
            if (lane == 0) red[OFF_B3] += acc_b3;
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int m = 16 * mt + grp + 8 * (c >> 1), nn = 8 * nt + 2 * tig + (c & 1);
                        red[OFF_W2 + m * HID + nn] += dw2_sum[mt][nt][c] + dw2[mt][nt][c];
                    }
        }
        __syncthreads();
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    float* dst = block_partials + (int64_t)blockIdx.x * DIFFUS_MLP_NPARAMS;
    for (int i = tid; i < DIFFUS_MLP_NPARAMS; i += TILE_M) dst[i] = red[i];
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(TMEM_COLS_BWD));
}

int mlp_bwd_tc_blocks(int64_t n) {
    int64_t tiles = (n + TILE_M - 1) / TILE_M;
    return (int)max((int64_t)1, min(tiles, (int64_t)148 * 2));     // 2 CTAs per SM (registers: 32 dW2 accumulators + the 32-wide epilogues)
}

cudaError_t launch_mlp_bwd_tc(const float* params, const float* x, const uint8_t* mask, const float* grad_out, int64_t n,
                              float out_scale, float* block_partials, int blocks, cudaStream_t st) {
    size_t smem = BT_SMEM_FLOATS * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(mlp_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(mlp_bwd_tc_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    mlp_bwd_tc_kernel<<<blocks, TILE_M, smem, st>>>(params, x, mask, grad_out, n, out_scale, block_partials);
    return cudaGetLastError();
}

}  // namespace diffus
