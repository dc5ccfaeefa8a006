// Probe: which shared-memory layouts does tcgen05.mma kind::tf32 accept for MN-major (transposed) operands?
// Host data A[128][128] (m, k) and B[96][128] (n, k), small integers (exact in tf32).
//   fill 0: no-swizzle MN-major   [group of 4 rows][k][4 floats]
//   fill 1: no-swizzle K-major    [group of 4 k][row][4 floats]
//   fill 2: 128B-swizzle tile     atom of 32 rows: [k][32 rows] (128 B per k), 16-byte chunk index XOR (k & 7)
//           read MN-major it is D = A B^T over k; read K-major it is the transposed problem
//           D'[k1][k2] = sum_{u<32} A[u][k1] B[u][k2]
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>

constexpr int M = 128, KTOT = 128, NMAX = 96;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout_type) {
    uint64_t d = (uint64_t)((addr >> 4) & 0x3fffu);
    d |= (uint64_t)((lbo >> 4) & 0x3fffu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3fffu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout_type << 61;
    return d;
}

struct Variant {
    const char* name;
    int fill, N;
    uint32_t a_lbo, a_sbo, b_lbo, b_sbo, layout_type, major;
    int ksteps;
    uint32_t kstep;
    int expect;      // 0: sum_k A[m][k] B[n][k] over 8*ksteps k's; 1: transposed problem over 8*ksteps units
};

__global__ void __launch_bounds__(128) probe(const float* A, const float* B, float* D, Variant v) {
    extern __shared__ __align__(1024) float smem[];
    float* a_t = smem;                         // 64 KB
    float* b_t = smem + M * KTOT;              // 48 KB
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int e = tid; e < (M + NMAX) * KTOT; e += 128) {
        const bool isb = e >= M * KTOT;
        const int ee = isb ? e - M * KTOT : e, rows = isb ? NMAX : M;
        const int r = ee / KTOT, k = ee % KTOT;
        float* t = isb ? b_t : a_t;
        const float val = isb ? B[ee] : A[ee];
        int idx;
        if (v.fill == 0) idx = (r >> 2) * KTOT * 4 + k * 4 + (r & 3);
        else if (v.fill == 1) idx = (k >> 2) * rows * 4 + r * 4 + (k & 3);
        else idx = (r >> 5) * KTOT * 32 + k * 32 + ((((r & 31) >> 2) ^ (k & 7)) << 2) + (r & 3);
        t[idx] = val;
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&slot)), "n"(128));
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
    const uint32_t tmem = slot;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | v.major | ((uint32_t)(v.N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    if (tid == 0) {
        uint32_t acc = 0;
        for (int ks = 0; ks < v.ksteps; ++ks) {
            uint64_t ad = smem_desc(smem_u32(a_t) + ks * v.kstep, v.a_lbo, v.a_sbo, v.layout_type);
            uint64_t bd = smem_desc(smem_u32(b_t) + ks * v.kstep, v.b_lbo, v.b_sbo, v.layout_type);
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem),
                         "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
            acc = 1;
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bar)) : "memory");
    }
    uint32_t done = 0;
    for (int spin = 0; spin < (1 << 22) && !done; ++spin)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;\n");
    for (int c0 = 0; c0 < v.N; c0 += 16) {
        uint32_t r[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                       "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        for (int j = 0; j < 16; ++j) D[tid * NMAX + c0 + j] = done ? __uint_as_float(r[j]) : -12345.f;
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(128));
}

int main() {
    static float hA[M * KTOT], hB[NMAX * KTOT], hD[M * NMAX];
    srand(1);
    for (auto& v : hA) v = (float)(rand() % 17 - 8);
    for (auto& v : hB) v = (float)(rand() % 13 - 6);
    float *A, *B, *D;
    cudaMalloc(&A, sizeof hA); cudaMalloc(&B, sizeof hB); cudaMalloc(&D, sizeof hD);
    cudaMemcpy(A, hA, sizeof hA, cudaMemcpyHostToDevice); cudaMemcpy(B, hB, sizeof hB, cudaMemcpyHostToDevice);
    size_t smem = (M + NMAX) * KTOT * sizeof(float) + 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const uint32_t MN = (1u << 15) | (1u << 16), G = KTOT * 16, ATOM = KTOT * 128;
    Variant vs[] = {
        {"K-major no-swizzle (control)", 1, 96, M * 16, 128, NMAX * 16, 128, 0, 0, KTOT / 8, 2 * M * 16, 0},
        {"MN-major no-swizzle lbo=128 sbo=G", 0, 80, 128, G, 128, G, 0, MN, KTOT / 8, 128, 0},
        {"MN-major no-swizzle lbo=G sbo=128", 0, 80, G, 128, G, 128, 0, MN, KTOT / 8, 128, 0},
        {"MN-major SW128 lbo=atom sbo=1024", 2, 96, ATOM, 1024, ATOM, 1024, 2, MN, KTOT / 8, 1024, 0},
        {"MN-major SW128 lbo=1024 sbo=atom", 2, 96, 1024, ATOM, 1024, ATOM, 2, MN, KTOT / 8, 1024, 0},
        {"K-major SW128 same tile (transposed problem)", 2, 96, 16, 1024, 16, 1024, 2, 0, 4, 32, 1},
    };
    for (auto& v : vs) {
        if (v.fill == 1) v.kstep = 0;   // per-operand K advance differs for A and B in this layout: handled below
        cudaMemset(D, 0, sizeof hD);
        Variant vv = v;
        if (v.fill == 1) { vv.ksteps = 1; }   // control: a single K=8 step
        probe<<<1, 128, smem>>>(A, B, D, vv);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: CUDA error %s\n", v.name, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(hD, D, sizeof hD, cudaMemcpyDeviceToHost);
        double maxerr = 0, maxabs = 0; int nz = 0;
        for (int m = 0; m < M; ++m)
            for (int n = 0; n < v.N; ++n) {
                double want = 0;
                if (v.expect == 0) for (int k = 0; k < vv.ksteps * 8; ++k) want += (double)hA[m * KTOT + k] * hB[n * KTOT + k];
                else for (int u = 0; u < vv.ksteps * 8; ++u) want += (double)hA[u * KTOT + m] * hB[u * KTOT + n];
                double got = hD[m * NMAX + n];
                maxerr = fmax(maxerr, fabs(got - want)); maxabs = fmax(maxabs, fabs(want)); nz += got != 0;
            }
        printf("%-46s max|err| %8.3f (max|want| %4.0f) nonzero %5d  D[0][0..3] = %.0f %.0f %.0f %.0f\n", v.name, maxerr, maxabs, nz,
               hD[0], hD[1], hD[2], hD[3]);
    }
    return 0;
}
