// Shared device helpers for the DiffUS B-mode renderer kernels (sm_100a).
//
// Column convention used by every kernel in this directory
// --------------------------------------------------------
// A ray has S samples; after the `start` crop there are Sout = S - start output columns
// c = 0..Sout-1, column c belonging to sample k = start + c.  Column c >= 1 owns the
// interface between samples c-1 and c with reflection coefficient
//     r_c = (Z_c - Z_{c-1}) / (Z_{c-1} + Z_c)                (reference src/renderer.py:33,65-68)
// and transfer matrix M_c = [[1 - 2 r_c^2, r_c], [-r_c, 1]]; M_0 = I.  With
//     P_c = M_0 M_1 ... M_c,     echo[c] = P_c[0][1] / P_c[1][1]
// the echo line equals what the reference obtains from one dense linear solve per
// truncation depth (src/renderer.py:367-457); see oracle/port.py::echo_closed_form.
//
// A warp owns a ray and walks it in passes of 512 columns: a gather phase with
// lane = consecutive sample (coalesced stores, few cache lines per load instruction)
// parks impedances in shared memory, a chunk phase with lane = CH consecutive columns does
// the sequential 2x2 products plus one warp-shuffle scan per 32*CH columns, and a tile phase
// writes the result back with lane = consecutive column.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <float.h>

#include "../../include/diffus_b200.h"

namespace diffus {

constexpr unsigned FULL = 0xffffffffu;

// Geometry of one warp pass: 32 lanes x CH consecutive columns per lane.  The forward uses
// CH = 16 (one warp scan per 512 columns); the backward keeps per-column prefixes in
// registers for its reverse sweep and uses CH = 8 to fit 128 registers (4 CTAs per SM) without spills.
template <int CH, int CK = 2>
struct Geo {
    static constexpr int CHUNK = CH;
    static constexpr int CKPT = CK;      // backward: columns between saved prefixes
    static constexpr int SEG = 32 * CH;
    // padded shared-memory slot: rows of CH+1 floats make the lane*CH+i pattern conflict free
    __device__ __forceinline__ static int pad(int i) { return i + i / CH; }
    static constexpr int ZBUF = SEG + 1 + (SEG + 1) / CH + 3;   // slots 0..SEG (slot 0 = sample c0-1)
    static constexpr int OBUF = SEG + SEG / CH + 3;             // slots 0..SEG-1
};
using FwdGeo = Geo<16>;
using BwdGeo = Geo<8, 2>;
// The forward saves its transfer-matrix carry at every PREFIX_STRIDE columns; the backward
// gathers PREFIX_STRIDE columns at a time and walks them as BWD_SUB sub-segments of BwdGeo::SEG.
constexpr int PREFIX_STRIDE = FwdGeo::SEG;
constexpr int BWD_SUB = PREFIX_STRIDE / BwdGeo::SEG;

struct M2 {          // [[a, b], [c, d]]
    float a, b, c, d;
};
__device__ __forceinline__ M2 m2_identity() { return M2{1.f, 0.f, 0.f, 1.f}; }
// The products below spell out their roundings (one multiply + one fused multiply-add per
// entry) so that every kernel that walks the same ray produces bit-identical prefixes,
// whatever the compiler would have contracted on its own.
__device__ __forceinline__ float mul_add2(float a, float b, float c, float d) {   // a*b + c*d
    return __fmaf_rn(a, b, __fmul_rn(c, d));
}
__device__ __forceinline__ M2 m2_mul(const M2& x, const M2& y) {
    return M2{mul_add2(x.a, y.a, x.b, y.c), mul_add2(x.a, y.b, x.b, y.d), mul_add2(x.c, y.a, x.d, y.c), mul_add2(x.c, y.b, x.d, y.d)};
}
// P * M(r),  M(r) = [[1 - 2 r^2, r], [-r, 1]]
__device__ __forceinline__ M2 m2_mul_interface(const M2& p, float r) {
    float q = __fmaf_rn(-2.f * r, r, 1.f);
    return M2{__fmaf_rn(p.a, q, -__fmul_rn(p.b, r)), __fmaf_rn(p.a, r, p.b), __fmaf_rn(p.c, q, -__fmul_rn(p.d, r)), __fmaf_rn(p.c, r, p.d)};
}
// X * M(r)^T
__device__ __forceinline__ M2 m2_mul_interface_t(const M2& x, float r) {
    float q = __fmaf_rn(-2.f * r, r, 1.f);
    return M2{mul_add2(x.a, q, x.b, r), __fmaf_rn(-x.a, r, x.b), mul_add2(x.c, q, x.d, r), __fmaf_rn(-x.c, r, x.d)};
}
__device__ __forceinline__ M2 m2_transpose(const M2& x) { return M2{x.a, x.c, x.b, x.d}; }
__device__ __forceinline__ M2 m2_add(const M2& x, const M2& y) { return M2{x.a + y.a, x.b + y.b, x.c + y.c, x.d + y.d}; }
__device__ __forceinline__ M2 m2_shfl_up(const M2& x, int d) {
    return M2{__shfl_up_sync(FULL, x.a, d), __shfl_up_sync(FULL, x.b, d), __shfl_up_sync(FULL, x.c, d), __shfl_up_sync(FULL, x.d, d)};
}
__device__ __forceinline__ M2 m2_shfl_down(const M2& x, int d) {
    return M2{__shfl_down_sync(FULL, x.a, d), __shfl_down_sync(FULL, x.b, d), __shfl_down_sync(FULL, x.c, d), __shfl_down_sync(FULL, x.d, d)};
}
__device__ __forceinline__ M2 m2_shfl(const M2& x, int src) {
    return M2{__shfl_sync(FULL, x.a, src), __shfl_sync(FULL, x.b, src), __shfl_sync(FULL, x.c, src), __shfl_sync(FULL, x.d, src)};
}

// nan_to_num(nan=0) of src/renderer.py:408 (+-inf -> +-FLT_MAX like torch's default)
__device__ __forceinline__ float nan_to_num(float e) {
    if (e != e) return 0.f;
    return fminf(fmaxf(e, -FLT_MAX), FLT_MAX);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
    return v;
}

// ---------------------------------------------------------------------------------------
// volume addressing: offset(i, j, k) = ox(i) + oy(j) + oz(k) in 32-bit element units, so the
// eight corners of a trilinear cell cost six per-axis terms and a handful of adds instead
// of eight full index computations (integer address math was 47 % of the forward's
// instructions in the first profile, profiles/r1_first_ncu.md).
// ---------------------------------------------------------------------------------------
struct VolumeView {
    const float* data;
    int D, H, W;          // extents along point components 0, 1, 2
    uint32_t sx, sy;      // LINEAR: H*W, W     BRICK: bricks-per-slab*32, bricks-per-row*32
};

constexpr int BRICK_I = 4, BRICK_J = 4, BRICK_K = 2;   // 32 floats = one 128-byte line

template <int LAYOUT>
__device__ __forceinline__ uint32_t axis_x(const VolumeView& v, int i) {
    return LAYOUT == DIFFUS_LAYOUT_LINEAR ? (uint32_t)i * v.sx : (uint32_t)(i >> 2) * v.sx + ((uint32_t)(i & 3) << 3);
}
template <int LAYOUT>
__device__ __forceinline__ uint32_t axis_y(const VolumeView& v, int j) {
    return LAYOUT == DIFFUS_LAYOUT_LINEAR ? (uint32_t)j * v.sy : (uint32_t)(j >> 2) * v.sy + ((uint32_t)(j & 3) << 1);
}
template <int LAYOUT>
__device__ __forceinline__ uint32_t axis_z(const VolumeView& v, int k) {
    return LAYOUT == DIFFUS_LAYOUT_LINEAR ? (uint32_t)k : ((uint32_t)(k >> 1) << 5) + (uint32_t)(k & 1);
}
template <int LAYOUT>
__device__ __forceinline__ uint32_t voxel_offset(const VolumeView& v, int i, int j, int k) {
    return axis_x<LAYOUT>(v, i) + axis_y<LAYOUT>(v, j) + axis_z<LAYOUT>(v, k);
}

// a / b without the IEEE slow path (<= 2 ulp): the reference's own float32 run is ~1e-5 of
// peak away from its float64 run, so correctly-rounded division buys nothing here
// One MUFU.RCP, no range fix-ups (__fdividef adds a scaling branch for |b| > 2^126, never met here).
__device__ __forceinline__ float fast_rcp(float b) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b));
    return r;
}
__device__ __forceinline__ float fast_div(float a, float b) { return a * fast_rcp(b); }
// echo of a prefix: P[0][1] / P[1][1] (same rounding in the forward and in the fused backward)
__device__ __forceinline__ float echo_of(float pb, float inv_pd) { return __fmul_rn(pb, inv_pd); }

// ---------------------------------------------------------------------------------------
// ray points:  p = source + k * direction   (src/renderer.py:119-124), cast to float32 (:751)
// The reference multiplies and adds in separate roundings; no FMA contraction here so the
// nearest-voxel indices are bit-identical.
// ---------------------------------------------------------------------------------------
template <bool POSE64>
struct RaySetup;

template <>
struct RaySetup<false> {
    float s[3], d[3];
    __device__ __forceinline__ void load(const void* src, const void* dir, int64_t pose, int64_t ray,
                                         int64_t n_rays, int64_t dir_pose_stride, int) {
        const float* sp = (const float*)src + pose * 3;
        const float* dp = (const float*)dir + pose * dir_pose_stride + ray * 3;
#pragma unroll
        for (int a = 0; a < 3; ++a) { s[a] = __ldg(sp + a); d[a] = __ldg(dp + a); }
    }
    __device__ __forceinline__ float coord(int a, int k) const {
        return __fadd_rn(s[a], __fmul_rn((float)k, d[a]));
    }
};

template <>
struct RaySetup<true> {
    double s[3], d[3];
    int product_f32;
    __device__ __forceinline__ void load(const void* src, const void* dir, int64_t pose, int64_t ray,
                                         int64_t n_rays, int64_t dir_pose_stride, int prod_f32) {
        const double* sp = (const double*)src + pose * 3;
        const double* dp = (const double*)dir + pose * dir_pose_stride + ray * 3;
#pragma unroll
        for (int a = 0; a < 3; ++a) { s[a] = __ldg(sp + a); d[a] = __ldg(dp + a); }
        product_f32 = prod_f32;
    }
    __device__ __forceinline__ float coord(int a, int k) const {
        double prod = product_f32 ? (double)__fmul_rn((float)k, (float)d[a]) : __dmul_rn((double)k, d[a]);
        return (float)__dadd_rn(s[a], prod);
    }
};

// ---------------------------------------------------------------------------------------
// samplers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ int nearest_index(float p, int n) {
    int i = __float2int_rn(p);                 // round half to even == torch.round; saturating
    return min(max(i, 0), n - 1);
}

struct TriCell {       // clamp-then-floor cell of a trilinear sample, grid_sample border semantics
    int i0[3], i1[3];
    float f[3];
    bool inside[3];    // derivative w.r.t. the coordinate is non-zero only strictly inside
};

__device__ __forceinline__ void tri_axis(float p, int n, int& i0, int& i1, float& f, bool& inside) {
    float hi = (float)(n - 1);
    inside = (p > 0.f) && (p < hi);
    float pc = fminf(fmaxf(p, 0.f), hi);
    float fl = floorf(pc);
    f = pc - fl;
    i0 = (int)fl;
    i1 = min(i0 + 1, n - 1);
}

template <int LAYOUT>
__device__ __forceinline__ void tri_offsets(const VolumeView& v, const TriCell& c, uint32_t off[8]) {
    uint32_t x0 = axis_x<LAYOUT>(v, c.i0[0]), x1 = axis_x<LAYOUT>(v, c.i1[0]);
    uint32_t y0 = axis_y<LAYOUT>(v, c.i0[1]), y1 = axis_y<LAYOUT>(v, c.i1[1]);
    uint32_t z0 = axis_z<LAYOUT>(v, c.i0[2]), z1 = axis_z<LAYOUT>(v, c.i1[2]);
    uint32_t a00 = x0 + y0, a01 = x0 + y1, a10 = x1 + y0, a11 = x1 + y1;
    off[0] = a00 + z0; off[1] = a00 + z1; off[2] = a01 + z0; off[3] = a01 + z1;
    off[4] = a10 + z0; off[5] = a10 + z1; off[6] = a11 + z0; off[7] = a11 + z1;
}

// value (and optionally the spatial gradient) of the border-clamped trilinear interpolant
template <bool GRAD>
__device__ __forceinline__ float tri_combine(const float z[8], const TriCell& c, float g[3]) {
    float fx = c.f[0], fy = c.f[1], fz = c.f[2];
    float gx = 1.f - fx, gy = 1.f - fy, gz = 1.f - fz;
    float c00 = z[0] * gz + z[1] * fz;   // (i0, j0)
    float c01 = z[2] * gz + z[3] * fz;   // (i0, j1)
    float c10 = z[4] * gz + z[5] * fz;   // (i1, j0)
    float c11 = z[6] * gz + z[7] * fz;   // (i1, j1)
    float c0 = c00 * gy + c01 * fy;
    float c1 = c10 * gy + c11 * fy;
    if (GRAD) {
        g[0] = c.inside[0] ? (c1 - c0) : 0.f;
        g[1] = c.inside[1] ? ((c01 - c00) * gx + (c11 - c10) * fx) : 0.f;
        float d00 = z[1] - z[0], d01 = z[3] - z[2], d10 = z[5] - z[4], d11 = z[7] - z[6];
        g[2] = c.inside[2] ? ((d00 * gy + d01 * fy) * gx + (d10 * gy + d11 * fy) * fx) : 0.f;
    }
    return c0 * gx + c1 * fx;
}

// A sample split into "issue the loads" and "combine", so a gather loop can keep the next
// tile's eight loads in flight while it combines the current one (software pipelining).
template <int SAMPLER, int LAYOUT>
struct Fetch {
    float z[SAMPLER == DIFFUS_SAMPLER_NEAREST ? 1 : 8];
    float f[3];
    bool inside[3];

    __device__ __forceinline__ void issue(const VolumeView& v, float p0, float p1, float p2) {
        if (SAMPLER == DIFFUS_SAMPLER_NEAREST) {
            int i = nearest_index(p0, v.D), j = nearest_index(p1, v.H), k = nearest_index(p2, v.W);
            z[0] = __ldg(v.data + voxel_offset<LAYOUT>(v, i, j, k));
        } else {
            TriCell c;
            tri_axis(p0, v.D, c.i0[0], c.i1[0], f[0], inside[0]);
            tri_axis(p1, v.H, c.i0[1], c.i1[1], f[1], inside[1]);
            tri_axis(p2, v.W, c.i0[2], c.i1[2], f[2], inside[2]);
            uint32_t off[8];
            tri_offsets<LAYOUT>(v, c, off);
#pragma unroll
            for (int q = 0; q < 8; ++q) z[q] = __ldg(v.data + off[q]);
        }
    }
    template <bool GRAD>
    __device__ __forceinline__ float finish(float g[3]) const {
        if (SAMPLER == DIFFUS_SAMPLER_NEAREST) {
            if (GRAD) { g[0] = g[1] = g[2] = 0.f; }
            return z[0];
        } else {
            TriCell c;
#pragma unroll
            for (int a = 0; a < 3; ++a) { c.f[a] = f[a]; c.inside[a] = inside[a]; }
            return tri_combine<GRAD>(z, c, g);
        }
    }
};

template <int SAMPLER, int LAYOUT, bool GRAD>
__device__ __forceinline__ float sample_volume(const VolumeView& v, float p0, float p1, float p2, float g[3]) {
    Fetch<SAMPLER, LAYOUT> fe;
    fe.issue(v, p0, p1, p2);
    return fe.template finish<GRAD>(g);
}

// kernel parameter block (host fills it from the C-ABI structs)
struct RenderParams {
    VolumeView vol;
    const void* sources;
    const void* directions;
    int64_t dir_pose_stride;
    int product_f32;
    int64_t n_poses, n_rays, total_rays;
    int S, start, Sout;
    int nprefix;                // saved prefixes per ray: ceil(Sout / PREFIX_STRIDE) - 1
    int att_slots;              // floats reserved at the start of dynamic smem for the attenuation table
    int att_slots_padded;       // same, for the backward's padded table
    float alpha;
    float* frame;
    float* seg_prefix;          // (total_rays, nprefix, 4) or null
    const float* median;        // (P) replacement for r_1 when start > 0, else null
    // backward
    const float* grad_frame;
    float* grad_volume;
    float* grad_src_partial;    // (total_rays, 3)
    float* grad_dir;            // (total_rays, 3)
    float* first_rbar;          // (total_rays) d loss / d (median-replaced r_1), start > 0
    // fused MSE loss (target != null): d loss / d frame = grad_scale * (frame - target)
    const float* target;        // (total_rays, Sout)
    float grad_scale;
    float* loss_partial;        // (total_rays) sum over the ray of (frame - target)^2, or null
};

}  // namespace diffus
