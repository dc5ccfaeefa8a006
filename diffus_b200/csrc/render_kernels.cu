// Fused per-ray march, forward and backward (sm_100a).  See common.cuh for the layout.
//
// forward  : replaces plot_beam_frame's numeric core (reference src/renderer.py:201-275)
// backward : what torch autograd derives from it, written as a reverse
//            affine scan of 2x2 matrices; per-ray gradient partials are written once
//            (no atomics) and only the volume gradient uses red.global.add.f32, mirroring
//            index_put_(accumulate=True).
#include "common.cuh"
#include "launch.h"

namespace diffus {

// ---------------------------------------------------------------------------------------
// chunk-phase building blocks shared by the render and the echo-only kernels
// ---------------------------------------------------------------------------------------

// exclusive prefix of the lanes' chunk products, left-multiplied by the segment carry
__device__ __forceinline__ M2 warp_exclusive_prefix(const M2& chunk_total, const M2& carry, int lane) {
    M2 inc = chunk_total;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        M2 left = m2_shfl_up(inc, d);
        if (lane >= d) inc = m2_mul(left, inc);
    }
    M2 exc = m2_shfl_up(inc, 1);
    if (lane == 0) exc = m2_identity();
    return m2_mul(carry, exc);
}

// Forward chunk phase for one segment.  `r[i]` must hold the coefficient of column
// c0 + lane*16 + i (0 for columns that do not exist).  Writes echo (NaN -> 0) to obuf and
// returns the updated carry (prefix through the last column of the segment).
__device__ __forceinline__ M2 forward_chunk(const float r[CHUNK], M2 carry, float* obuf, int lane) {
    M2 T = m2_identity();
#pragma unroll
    for (int i = 0; i < CHUNK; ++i) T = m2_mul_interface(T, r[i]);
    M2 P = warp_exclusive_prefix(T, carry, lane);
#pragma unroll
    for (int i = 0; i < CHUNK; ++i) {
        P = m2_mul_interface(P, r[i]);
        obuf[pad(lane * CHUNK + i)] = nan_to_num(P.b / P.d);
    }
    return m2_shfl(P, 31);
}

// Backward chunk phase for one segment.
//   r[i]        coefficient of column c0 + lane*16 + i (0 where the column does not exist)
//   gbuf        in: d loss / d echo per column (0 where none); out: d loss / d r per column
//   carry       forward prefix P through the column before the segment
//   vin         adjoint flowing into the segment's last column from later segments
// returns the adjoint flowing out of the segment's first column (into the previous segment)
__device__ __forceinline__ M2 backward_chunk(const float r[CHUNK], const M2& carry, const M2& vin,
                                             float* gbuf, int lane) {
    // local chunk product and true prefixes
    M2 T = m2_identity();
#pragma unroll
    for (int i = 0; i < CHUNK; ++i) T = m2_mul_interface(T, r[i]);
    M2 P = warp_exclusive_prefix(T, carry, lane);
    M2 Pst[CHUNK];                 // true prefix BEFORE column i of the chunk
    M2 G = m2_identity();          // local inclusive prefix
    M2 B = M2{0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < CHUNK; ++i) {
        Pst[i] = P;
        P = m2_mul_interface(P, r[i]);
        G = m2_mul_interface(G, r[i]);
        float e = P.b / P.d;
        float ge = gbuf[pad(lane * CHUNK + i)];
        if (!(fabsf(e) <= FLT_MAX)) ge = 0.f;          // nan_to_num passes no gradient at NaN/inf
        float inv = 1.f / P.d;
        float da = ge * inv, db = -ge * e * inv;       // D = [[0, da], [0, db]]
        B.a += da * G.b; B.b += da * G.d; B.c += db * G.b; B.d += db * G.d;
    }
    // suffix scan of the affine maps X -> X * A + B, A = T^T
    M2 A = m2_transpose(T);
    M2 As = A, Bs = B;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        M2 An = m2_shfl_down(As, d), Bn = m2_shfl_down(Bs, d);
        if (lane + d < 32) {
            Bs = m2_add(m2_mul(Bn, As), Bs);
            As = m2_mul(An, As);
        }
    }
    // adjoint entering this lane's last column: maps of lanes lane+1..31 applied to vin
    M2 An = m2_shfl_down(As, 1), Bn = m2_shfl_down(Bs, 1);
    M2 V = (lane == 31) ? vin : m2_add(m2_mul(vin, An), Bn);
    M2 vout = m2_add(m2_mul(vin, As), Bs);             // valid on lane 0
#pragma unroll
    for (int i = CHUNK - 1; i >= 0; --i) {
        M2 Pc = m2_mul_interface(Pst[i], r[i]);
        float e = Pc.b / Pc.d;
        float ge = gbuf[pad(lane * CHUNK + i)];
        if (!(fabsf(e) <= FLT_MAX)) ge = 0.f;
        float inv = 1.f / Pc.d;
        M2 Pbar = V;
        Pbar.b += ge * inv;
        Pbar.d += -ge * e * inv;
        const M2& Q = Pst[i];
        float ma = Q.a * Pbar.a + Q.c * Pbar.c;
        float mb = Q.a * Pbar.b + Q.c * Pbar.d;
        float mc = Q.b * Pbar.a + Q.d * Pbar.c;
        float rbar = -4.f * r[i] * ma + mb - mc;
        gbuf[pad(lane * CHUNK + i)] = (rbar == rbar) ? rbar : 0.f;
        V = m2_mul_interface_t(Pbar, r[i]);
    }
    return m2_shfl(vout, 0);
}

// coefficient of the interface owned by column c from the two impedances around it
__device__ __forceinline__ float reflection(float z_prev, float z_cur) { return (z_cur - z_prev) / (z_prev + z_cur); }

// ---------------------------------------------------------------------------------------
// forward render
// ---------------------------------------------------------------------------------------
template <int SAMPLER, int LAYOUT, bool POSE64>
__global__ void __launch_bounds__(128) render_fwd_kernel(const RenderParams p) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t ray = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    if (ray >= p.total_rays) return;
    float* zbuf = smem + warp * (ZBUF + OBUF);
    float* obuf = zbuf + ZBUF;
    const int64_t pose = ray / p.n_rays;
    RaySetup<POSE64> rs;
    rs.load(p.sources, p.directions, pose, ray - pose * p.n_rays, p.n_rays, p.dir_pose_stride, p.product_f32);
    const float med = p.median ? __ldg(p.median + pose) : 0.f;
    float* out = p.frame + ray * (int64_t)p.Sout;

    M2 carry = m2_identity();
    for (int s = 0; s < p.nseg; ++s) {
        const int c0 = s * SEG;
        const int ncol = min(SEG, p.Sout - c0);
        // gather phase: lane = consecutive sample
        const int ntile = (ncol + 31) >> 5;
#pragma unroll 4
        for (int t = 0; t < ntile; ++t) {
            int idx = t * 32 + lane;
            if (idx < ncol) {
                int k = p.start + c0 + idx;
                float g[3];
                float z = sample_volume<SAMPLER, LAYOUT, false>(p.vol, rs.coord(0, k), rs.coord(1, k), rs.coord(2, k), g);
                zbuf[pad(idx + 1)] = z;
            }
        }
        __syncwarp();
        // chunk phase: lane = 16 consecutive columns
        float r[CHUNK];
        {
            int cl = lane * CHUNK;
            float zp = zbuf[pad(cl)];
#pragma unroll
            for (int i = 0; i < CHUNK; ++i) {
                float zc = zbuf[pad(cl + i + 1)];
                int c = c0 + cl + i;
                float ri = reflection(zp, zc);
                if (c == 1 && p.median) ri = med;
                r[i] = (c >= 1 && cl + i < ncol) ? ri : 0.f;
                zp = zc;
            }
        }
        carry = forward_chunk(r, carry, obuf, lane);
        if (p.seg_prefix && s + 1 < p.nseg && lane == 0) {
            float4* sp = (float4*)(p.seg_prefix + (ray * (p.nseg - 1) + s) * 4);
            *sp = make_float4(carry.a, carry.b, carry.c, carry.d);
        }
        __syncwarp();
        // tile phase: attenuate and write, lane = consecutive column
#pragma unroll 4
        for (int t = 0; t < ntile; ++t) {
            int idx = t * 32 + lane;
            if (idx < ncol) {
                int c = c0 + idx;
                out[c] = obuf[pad(idx)] * expf(-p.alpha * (float)c);
            }
        }
        if (lane == 0) zbuf[pad(0)] = zbuf[pad(SEG)];   // sample c0+SEG-1 becomes the next segment's left neighbour
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------
// backward render
// ---------------------------------------------------------------------------------------
constexpr int BWD_SMEM_PER_WARP = ZBUF + OBUF + 3 * SEG;

template <int SAMPLER, bool POSE64>
__device__ __forceinline__ void scatter_volume_grad(const RenderParams& p, const RaySetup<POSE64>& rs, int k, float zbar) {
    // gradient volume is always LINEAR (it is handed back to torch / the MLP backward)
    float p0 = rs.coord(0, k), p1 = rs.coord(1, k), p2 = rs.coord(2, k);
    if (SAMPLER == DIFFUS_SAMPLER_NEAREST) {
        int i = nearest_index(p0, p.vol.D), j = nearest_index(p1, p.vol.H), kk = nearest_index(p2, p.vol.W);
        atomicAdd(p.grad_volume + ((int64_t)i * p.vol.H + j) * p.vol.W + kk, zbar);
    } else {
        TriCell c;
        tri_axis(p0, p.vol.D, c.i0[0], c.i1[0], c.f[0], c.inside[0]);
        tri_axis(p1, p.vol.H, c.i0[1], c.i1[1], c.f[1], c.inside[1]);
        tri_axis(p2, p.vol.W, c.i0[2], c.i1[2], c.f[2], c.inside[2]);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            int i = (q & 4) ? c.i1[0] : c.i0[0];
            int j = (q & 2) ? c.i1[1] : c.i0[1];
            int kk = (q & 1) ? c.i1[2] : c.i0[2];
            float w = ((q & 4) ? c.f[0] : 1.f - c.f[0]) * ((q & 2) ? c.f[1] : 1.f - c.f[1]) * ((q & 1) ? c.f[2] : 1.f - c.f[2]);
            if (w != 0.f) atomicAdd(p.grad_volume + ((int64_t)i * p.vol.H + j) * p.vol.W + kk, w * zbar);
        }
    }
}

template <int SAMPLER, int LAYOUT, bool POSE64, bool POSE_GRAD, bool VOL_GRAD>
__global__ void __launch_bounds__(128) render_bwd_kernel(const RenderParams p) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t ray = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    if (ray >= p.total_rays) return;
    float* zbuf = smem + warp * BWD_SMEM_PER_WARP;
    float* gbuf = zbuf + ZBUF;
    float* dz = gbuf + OBUF;                 // [3][SEG] spatial gradient of Z at each sample
    const int64_t pose = ray / p.n_rays;
    RaySetup<POSE64> rs;
    rs.load(p.sources, p.directions, pose, ray - pose * p.n_rays, p.n_rays, p.dir_pose_stride, p.product_f32);
    const float med = p.median ? __ldg(p.median + pose) : 0.f;
    const float* gout = p.grad_frame + ray * (int64_t)p.Sout;

    M2 vin = M2{0.f, 0.f, 0.f, 0.f};
    float carry_rbar = 0.f, carry_z = 0.f;   // column c0+SEG of the later segment: its d loss/d r and its impedance
    float acc_s[3] = {0.f, 0.f, 0.f}, acc_d[3] = {0.f, 0.f, 0.f};

    for (int s = p.nseg - 1; s >= 0; --s) {
        const int c0 = s * SEG;
        const int ncol = min(SEG, p.Sout - c0);
        const int ntile = (ncol + 31) >> 5;
        // gather phase (re-gather: nothing but the segment prefixes is saved by the forward)
#pragma unroll 2
        for (int t = 0; t < ntile; ++t) {
            int idx = t * 32 + lane;
            if (idx < ncol) {
                int c = c0 + idx, k = p.start + c;
                float g[3];
                float z = sample_volume<SAMPLER, LAYOUT, POSE_GRAD>(p.vol, rs.coord(0, k), rs.coord(1, k), rs.coord(2, k), g);
                zbuf[pad(idx + 1)] = z;
                if (POSE_GRAD) { dz[idx] = g[0]; dz[SEG + idx] = g[1]; dz[2 * SEG + idx] = g[2]; }
                gbuf[pad(idx)] = __ldg(gout + c) * expf(-p.alpha * (float)c);
            }
        }
        if (s > 0 && lane == 0) {            // left neighbour of the segment's first column
            int k = p.start + c0 - 1;
            float g[3];
            zbuf[pad(0)] = sample_volume<SAMPLER, LAYOUT, false>(p.vol, rs.coord(0, k), rs.coord(1, k), rs.coord(2, k), g);
        }
        // columns that do not exist carry no gradient
        for (int idx = ncol + lane; idx < SEG; idx += 32) gbuf[pad(idx)] = 0.f;
        __syncwarp();

        float r[CHUNK];
        {
            int cl = lane * CHUNK;
            float zp = zbuf[pad(cl)];
#pragma unroll
            for (int i = 0; i < CHUNK; ++i) {
                float zc = zbuf[pad(cl + i + 1)];
                int c = c0 + cl + i;
                float ri = reflection(zp, zc);
                if (c == 1 && p.median) ri = med;
                r[i] = (c >= 1 && cl + i < ncol) ? ri : 0.f;
                zp = zc;
            }
        }
        M2 carry = m2_identity();
        if (s > 0) {
            float4 c4 = __ldg((const float4*)(p.seg_prefix + (ray * (p.nseg - 1) + (s - 1)) * 4));
            carry = M2{c4.x, c4.y, c4.z, c4.w};
        }
        vin = backward_chunk(r, carry, vin, gbuf, lane);
        __syncwarp();

        // tile phase: d loss / d Z per sample, then pose partials and the volume scatter
        float next_carry_rbar = gbuf[pad(0)], next_carry_z = zbuf[pad(1)];
        for (int t = 0; t < ntile; ++t) {
            int idx = t * 32 + lane;
            if (idx < ncol) {
                int c = c0 + idx;
                float zc = zbuf[pad(idx + 1)];
                float zbar = 0.f;
                // as the right-hand impedance of its own column's interface
                if (c >= 1 && !(c == 1 && p.median)) {
                    float zp = zbuf[pad(idx)];
                    float sum = zp + zc;
                    zbar += gbuf[pad(idx)] * (2.f * zp / (sum * sum));
                }
                // as the left-hand impedance of the next column's interface
                if (c + 1 < p.Sout && !(c == 0 && p.median)) {
                    float zn, rb;
                    if (idx + 1 < ncol) { zn = zbuf[pad(idx + 2)]; rb = gbuf[pad(idx + 1)]; }
                    else { zn = carry_z; rb = carry_rbar; }
                    float sum = zc + zn;
                    zbar -= rb * (2.f * zn / (sum * sum));
                }
                if (!(zbar == zbar)) zbar = 0.f;
                int k = p.start + c;
                if (POSE_GRAD) {
                    float kf = (float)k;
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        float ga = zbar * dz[a * SEG + idx];
                        acc_s[a] += ga;
                        acc_d[a] += kf * ga;
                    }
                }
                if (VOL_GRAD && zbar != 0.f) scatter_volume_grad<SAMPLER, POSE64>(p, rs, k, zbar);
            }
        }
        if (p.first_rbar && s == 0 && lane == 0) p.first_rbar[ray] = gbuf[pad(1)];
        carry_rbar = __shfl_sync(FULL, next_carry_rbar, 0);
        carry_z = __shfl_sync(FULL, next_carry_z, 0);
        __syncwarp();
    }
    if (POSE_GRAD) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            float ss = warp_sum(acc_s[a]), dd = warp_sum(acc_d[a]);
            if (lane == 0) {
                p.grad_src_partial[ray * 3 + a] = ss;
                p.grad_dir[ray * 3 + a] = dd;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// echo-only kernels: compute_echo_traces on explicit coefficients (src/renderer.py:439-457)
// refl (B, N) -> echo (B, N+1); column c >= 1 uses refl[c-1]
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) echo_fwd_kernel(const float* __restrict__ refl, int64_t n_rays, int N,
                                                       float* __restrict__ echo) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t ray = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    if (ray >= n_rays) return;
    float* rbuf = smem + warp * (2 * OBUF);
    float* obuf = rbuf + OBUF;
    const int Sout = N + 1, nseg = (Sout + SEG - 1) / SEG;
    const float* rin = refl + ray * (int64_t)N;
    float* out = echo + ray * (int64_t)Sout;
    M2 carry = m2_identity();
    for (int s = 0; s < nseg; ++s) {
        const int c0 = s * SEG, ncol = min(SEG, Sout - c0);
        for (int idx = lane; idx < SEG; idx += 32) {
            int c = c0 + idx;
            rbuf[pad(idx)] = (c >= 1 && idx < ncol) ? __ldg(rin + c - 1) : 0.f;
        }
        __syncwarp();
        float r[CHUNK];
#pragma unroll
        for (int i = 0; i < CHUNK; ++i) r[i] = rbuf[pad(lane * CHUNK + i)];
        carry = forward_chunk(r, carry, obuf, lane);
        __syncwarp();
        for (int idx = lane; idx < ncol; idx += 32) out[c0 + idx] = obuf[pad(idx)];
        __syncwarp();
    }
}

// two passes over the segments: forward to collect the prefixes, then the reverse scan
__global__ void __launch_bounds__(128) echo_bwd_kernel(const float* __restrict__ refl, const float* __restrict__ grad_echo,
                                                       int64_t n_rays, int N, float* __restrict__ grad_refl) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t ray = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    if (ray >= n_rays) return;
    const int Sout = N + 1, nseg = (Sout + SEG - 1) / SEG;
    float* rbuf = smem + warp * (2 * OBUF + 4 * nseg);
    float* gbuf = rbuf + OBUF;
    float* prefix = gbuf + OBUF;             // carry entering segment s, 4 floats each
    const float* rin = refl + ray * (int64_t)N;
    const float* gin = grad_echo + ray * (int64_t)Sout;
    float* gout = grad_refl + ray * (int64_t)N;
    M2 carry = m2_identity();
    for (int s = 0; s < nseg; ++s) {
        const int c0 = s * SEG, ncol = min(SEG, Sout - c0);
        if (lane == 0) { prefix[4 * s] = carry.a; prefix[4 * s + 1] = carry.b; prefix[4 * s + 2] = carry.c; prefix[4 * s + 3] = carry.d; }
        if (s + 1 == nseg) break;
        for (int idx = lane; idx < SEG; idx += 32) {
            int c = c0 + idx;
            rbuf[pad(idx)] = (c >= 1 && idx < ncol) ? __ldg(rin + c - 1) : 0.f;
        }
        __syncwarp();
        M2 T = m2_identity();
#pragma unroll
        for (int i = 0; i < CHUNK; ++i) T = m2_mul_interface(T, rbuf[pad(lane * CHUNK + i)]);
        M2 P = warp_exclusive_prefix(T, carry, lane);
        carry = m2_shfl(m2_mul(P, T), 31);
        __syncwarp();
    }
    __syncwarp();
    M2 vin = M2{0.f, 0.f, 0.f, 0.f};
    for (int s = nseg - 1; s >= 0; --s) {
        const int c0 = s * SEG, ncol = min(SEG, Sout - c0);
        for (int idx = lane; idx < SEG; idx += 32) {
            int c = c0 + idx;
            bool ok = idx < ncol;
            rbuf[pad(idx)] = (c >= 1 && ok) ? __ldg(rin + c - 1) : 0.f;
            gbuf[pad(idx)] = ok ? __ldg(gin + c) : 0.f;
        }
        __syncwarp();
        float r[CHUNK];
#pragma unroll
        for (int i = 0; i < CHUNK; ++i) r[i] = rbuf[pad(lane * CHUNK + i)];
        M2 cs = M2{prefix[4 * s], prefix[4 * s + 1], prefix[4 * s + 2], prefix[4 * s + 3]};
        vin = backward_chunk(r, cs, vin, gbuf, lane);
        __syncwarp();
        for (int idx = lane; idx < ncol; idx += 32) {
            int c = c0 + idx;
            if (c >= 1) gout[c - 1] = gbuf[pad(idx)];
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------
static int warps_per_block(int64_t total_rays) {
    // small problems: one ray per CTA so a single frame still spreads over the SMs
    if (total_rays >= 4 * 148 * 4) return 4;
    if (total_rays >= 2 * 148 * 2) return 2;
    return 1;
}

template <int SAMPLER, int LAYOUT, bool POSE64>
static cudaError_t launch_fwd_t(const RenderParams& p, cudaStream_t st) {
    int wpb = warps_per_block(p.total_rays);
    size_t smem = (size_t)wpb * (ZBUF + OBUF) * sizeof(float);
    int64_t grid = (p.total_rays + wpb - 1) / wpb;
    render_fwd_kernel<SAMPLER, LAYOUT, POSE64><<<(unsigned)grid, wpb * 32, smem, st>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_render_fwd(const RenderParams& p, int sampler, int layout, int pose64, cudaStream_t st) {
#define DIFFUS_FWD_CASE(S, L, P64) \
    if (sampler == S && layout == L && pose64 == (P64 ? 1 : 0)) return launch_fwd_t<S, L, P64>(p, st);
    DIFFUS_FWD_CASE(DIFFUS_SAMPLER_NEAREST, DIFFUS_LAYOUT_LINEAR, false)
    DIFFUS_FWD_CASE(DIFFUS_SAMPLER_NEAREST, DIFFUS_LAYOUT_LINEAR, true)
    DIFFUS_FWD_CASE(DIFFUS_SAMPLER_TRILINEAR, DIFFUS_LAYOUT_LINEAR, false)
    DIFFUS_FWD_CASE(DIFFUS_SAMPLER_TRILINEAR, DIFFUS_LAYOUT_LINEAR, true)
    DIFFUS_FWD_CASE(DIFFUS_SAMPLER_NEAREST, DIFFUS_LAYOUT_BRICK, false)
    DIFFUS_FWD_CASE(DIFFUS_SAMPLER_NEAREST, DIFFUS_LAYOUT_BRICK, true)
    DIFFUS_FWD_CASE(DIFFUS_SAMPLER_TRILINEAR, DIFFUS_LAYOUT_BRICK, false)
    DIFFUS_FWD_CASE(DIFFUS_SAMPLER_TRILINEAR, DIFFUS_LAYOUT_BRICK, true)
#undef DIFFUS_FWD_CASE
    return cudaErrorInvalidValue;
}

template <int SAMPLER, int LAYOUT, bool POSE64, bool PG, bool VG>
static cudaError_t launch_bwd_t(const RenderParams& p, cudaStream_t st) {
    int wpb = warps_per_block(p.total_rays);
    size_t smem = (size_t)wpb * BWD_SMEM_PER_WARP * sizeof(float);
    int64_t grid = (p.total_rays + wpb - 1) / wpb;
    render_bwd_kernel<SAMPLER, LAYOUT, POSE64, PG, VG><<<(unsigned)grid, wpb * 32, smem, st>>>(p);
    return cudaGetLastError();
}

template <int SAMPLER, int LAYOUT, bool POSE64>
static cudaError_t launch_bwd_g(const RenderParams& p, bool pg, bool vg, cudaStream_t st) {
    if (SAMPLER == DIFFUS_SAMPLER_NEAREST) pg = false;     // no pose gradient exists (round+long cuts the graph)
    if (pg && vg) return launch_bwd_t<SAMPLER, LAYOUT, POSE64, SAMPLER == DIFFUS_SAMPLER_TRILINEAR, true>(p, st);
    if (pg) return launch_bwd_t<SAMPLER, LAYOUT, POSE64, SAMPLER == DIFFUS_SAMPLER_TRILINEAR, false>(p, st);
    return launch_bwd_t<SAMPLER, LAYOUT, POSE64, false, true>(p, st);
}

cudaError_t launch_render_bwd(const RenderParams& p, int sampler, int layout, int pose64, bool pose_grad, bool vol_grad,
                              cudaStream_t st) {
#define DIFFUS_BWD_CASE(S, L, P64) \
    if (sampler == S && layout == L && pose64 == (P64 ? 1 : 0)) return launch_bwd_g<S, L, P64>(p, pose_grad, vol_grad, st);
    DIFFUS_BWD_CASE(DIFFUS_SAMPLER_NEAREST, DIFFUS_LAYOUT_LINEAR, false)
    DIFFUS_BWD_CASE(DIFFUS_SAMPLER_NEAREST, DIFFUS_LAYOUT_LINEAR, true)
    DIFFUS_BWD_CASE(DIFFUS_SAMPLER_TRILINEAR, DIFFUS_LAYOUT_LINEAR, false)
    DIFFUS_BWD_CASE(DIFFUS_SAMPLER_TRILINEAR, DIFFUS_LAYOUT_LINEAR, true)
    DIFFUS_BWD_CASE(DIFFUS_SAMPLER_NEAREST, DIFFUS_LAYOUT_BRICK, false)
    DIFFUS_BWD_CASE(DIFFUS_SAMPLER_NEAREST, DIFFUS_LAYOUT_BRICK, true)
    DIFFUS_BWD_CASE(DIFFUS_SAMPLER_TRILINEAR, DIFFUS_LAYOUT_BRICK, false)
    DIFFUS_BWD_CASE(DIFFUS_SAMPLER_TRILINEAR, DIFFUS_LAYOUT_BRICK, true)
#undef DIFFUS_BWD_CASE
    return cudaErrorInvalidValue;
}

cudaError_t launch_echo_fwd(const float* refl, int64_t n_rays, int N, float* echo, cudaStream_t st) {
    int wpb = warps_per_block(n_rays);
    size_t smem = (size_t)wpb * 2 * OBUF * sizeof(float);
    echo_fwd_kernel<<<(unsigned)((n_rays + wpb - 1) / wpb), wpb * 32, smem, st>>>(refl, n_rays, N, echo);
    return cudaGetLastError();
}

cudaError_t launch_echo_bwd(const float* refl, const float* grad_echo, int64_t n_rays, int N, float* grad_refl,
                            cudaStream_t st) {
    int wpb = warps_per_block(n_rays);
    int nseg = (N + 1 + SEG - 1) / SEG;
    size_t smem = (size_t)wpb * (2 * OBUF + 4 * nseg) * sizeof(float);
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(echo_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    echo_bwd_kernel<<<(unsigned)((n_rays + wpb - 1) / wpb), wpb * 32, smem, st>>>(refl, grad_echo, n_rays, N, grad_refl);
    return cudaGetLastError();
}

}  // namespace diffus
