// norm.cu -- BatchNorm family on NHWC fp32: per-channel statistics, finalize (+running stats),
// affine+activation apply, and the two-pass backward (reduce, apply).  All HBM-bound: coalesced
// 128-bit loads along the channel axis, fp64 accumulation of the per-channel sums.
#include "common.cuh"
#include <stdlib.h>

namespace saunet {

// ---- generic per-channel reduction over pixels ------------------------------------------------
// Thread (r, l): l indexes a VEC-wide channel group, r a pixel row inside the block.  Each op adds
// NV quantities per channel.  Result accumulated with double atomics into out[NV][C].
template <int VEC, typename Op>
__global__ void __launch_bounds__(256) chan_reduce_kernel(Op op, int C, long long npix, double* __restrict__ out, long long vstride) {
    constexpr int NV = Op::NV;
    __shared__ double red[NV * VEC * 256];
    const int L = C / VEC;                       // channel groups per pixel
    const int lanes = L < 256 ? L : 256;
    const int rows = 256 / lanes;
    const int tid = threadIdx.x;
    const int l = tid % lanes, r = tid / lanes;
    const bool active = r < rows;
    for (int l0 = 0; l0 < L; l0 += lanes) {
        const int lg = l0 + l;
        double acc[NV][VEC];
#pragma unroll
        for (int v = 0; v < NV; ++v)
#pragma unroll
            for (int e = 0; e < VEC; ++e) acc[v][e] = 0.0;
        if (active && lg < L) {
            for (long long p = (long long)blockIdx.x * rows + r; p < npix; p += (long long)gridDim.x * rows)
                op(p, lg * VEC, acc);
        }
        if (rows > 1) {
            __syncthreads();
#pragma unroll
            for (int v = 0; v < NV; ++v)
#pragma unroll
                for (int e = 0; e < VEC; ++e) red[(v * VEC + e) * 256 + tid] = acc[v][e];
            __syncthreads();
            if (r == 0 && lg < L) {
                for (int rr = 1; rr < rows; ++rr)
#pragma unroll
                    for (int v = 0; v < NV; ++v)
#pragma unroll
                        for (int e = 0; e < VEC; ++e) acc[v][e] += red[(v * VEC + e) * 256 + rr * lanes + l];
            }
        }
        if (r == 0 && lg < L) {
#pragma unroll
            for (int v = 0; v < NV; ++v)
#pragma unroll
                for (int e = 0; e < VEC; ++e) atomicAdd(out + (size_t)v * vstride + lg * VEC + e, acc[v][e]);
        }
    }
}

template <int VEC> struct Vec;
template <> struct Vec<1> {
    float v[1];
    __device__ static Vec load(const float* p) { Vec r; r.v[0] = __ldg(p); return r; }
    __device__ void store(float* p) const { p[0] = v[0]; }
};
template <> struct Vec<4> {
    float v[4];
    __device__ static Vec load(const float* p) {
        float4 t = __ldg(reinterpret_cast<const float4*>(p)); Vec r; r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w; return r;
    }
    __device__ void store(float* p) const { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
};

template <int VEC>
struct StatsOp {
    static constexpr int NV = 2;
    const float* x; int ld;
    __device__ void operator()(long long p, int c, double (&acc)[2][VEC]) const {
        Vec<VEC> v = Vec<VEC>::load(x + (size_t)p * ld + c);
#pragma unroll
        for (int e = 0; e < VEC; ++e) { double t = v.v[e]; acc[0][e] += t; acc[1][e] += t * t; }
    }
};

template <int VEC>
struct BnBwdOp {
    static constexpr int NV = 2;
    const float* dy; int dy_ld; const float* x; int x_ld; const float* out; int out_ld;
    const float* state; int C; int act;
    __device__ void operator()(long long p, int c, double (&acc)[2][VEC]) const {
        Vec<VEC> g = Vec<VEC>::load(dy + (size_t)p * dy_ld + c);
        Vec<VEC> xv = Vec<VEC>::load(x + (size_t)p * x_ld + c);
        Vec<VEC> ov;
        if (act == SAUNET_ACT_RELU && out) ov = Vec<VEC>::load(out + (size_t)p * out_ld + c);
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            const float mean = state[2 * C + c + e], invstd = state[3 * C + c + e];
            float gg = g.v[e];
            if (act == SAUNET_ACT_RELU) {
                float z = out ? ov.v[e] : fmaf(xv.v[e], state[c + e], state[C + c + e]);
                if (!(z > 0.f)) gg = 0.f;
            }
            float xh = (xv.v[e] - mean) * invstd;
            acc[0][e] += (double)gg; acc[1][e] += (double)gg * (double)xh;
        }
    }
};

template <typename Op1, typename Op4>
static int launch_reduce(const Op1& o1, const Op4& o4, bool vec, int C, long long npix, double* out, long long vstride, cudaStream_t st, const char* name) {
    int L = vec ? C / 4 : C;
    int lanes = L < 256 ? L : 256;
    int rows = 256 / lanes;
    long long want = (npix + rows - 1) / rows;
    // keep >= 8 pixels per thread when possible, cap at 8 CTAs/SM
    long long cap = (long long)kNumSMs * 8;
    long long blocks = want / 8; if (blocks < 1) blocks = 1; if (blocks > cap) blocks = cap;
    if (vec) chan_reduce_kernel<4, Op4><<<(int)blocks, 256, 0, st>>>(o4, C, npix, out, vstride);
    else chan_reduce_kernel<1, Op1><<<(int)blocks, 256, 0, st>>>(o1, C, npix, out, vstride);
    SAUNET_CHECK_LAUNCH(name);
    return SAUNET_OK;
}

// ---- finalize -----------------------------------------------------------------------------------
__global__ void bn_finalize_kernel(const double* __restrict__ sum, const double* __restrict__ sumsq, double count,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, float* running_mean,
                                   float* running_var, float momentum, float eps, int training, int C, float* __restrict__ state) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double mean, var;
    if (training) {
        mean = sum[c] / count;
        var = sumsq[c] / count - mean * mean;
        if (var < 0.0) var = 0.0;
        if (running_mean) {
            double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
            running_mean[c] = (float)((1.0 - (double)momentum) * (double)running_mean[c] + (double)momentum * mean);
            running_var[c] = (float)((1.0 - (double)momentum) * (double)running_var[c] + (double)momentum * unbiased);
        }
    } else {
        mean = running_mean[c]; var = running_var[c];
    }
    float invstd = (float)(1.0 / sqrt(var + (double)eps));
    float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
    float scale = g * invstd;
    state[c] = scale;
    state[C + c] = b - (float)mean * scale;
    state[2 * C + c] = (float)mean;
    state[3 * C + c] = invstd;
}

// ---- apply ---------------------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(256) affine_act_kernel(const float* __restrict__ x, int x_ld, const float* __restrict__ scale,
                                                         const float* __restrict__ shift, const float* __restrict__ res, int r_ld,
                                                         float* __restrict__ y, int y_ld, int C, long long npix, int act) {
    const int L = C / VEC;
    const long long n = npix * L;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
        long long p = idx / L; int c = (int)(idx - p * L) * VEC;
        Vec<VEC> v = Vec<VEC>::load(x + (size_t)p * x_ld + c);
        Vec<VEC> rv;
        if (res) rv = Vec<VEC>::load(res + (size_t)p * r_ld + c);
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            float t = scale ? fmaf(v.v[e], scale[c + e], shift[c + e]) : v.v[e];
            if (res) t += rv.v[e];
            v.v[e] = apply_act(t, act);
        }
        v.store(y + (size_t)p * y_ld + c);
    }
}

template <int VEC>
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float* __restrict__ dy, int dy_ld, const float* __restrict__ x, int x_ld,
                                                           const float* __restrict__ out, int out_ld, const float* __restrict__ state,
                                                           const float* __restrict__ gamma, const double* __restrict__ red, int C,
                                                           long long npix, int act, int training, float* dx, int dx_ld, int dx_acc,
                                                           float* dres, int dres_ld, int dres_acc, float* dgamma, float* dbeta) {
    const int L = C / VEC;
    const long long n = npix * L;
    const double inv_n = 1.0 / (double)npix;
    if (blockIdx.x == 0 && dgamma) {
        for (int c = threadIdx.x; c < C; c += blockDim.x) { dbeta[c] += (float)red[c]; dgamma[c] += (float)red[C + c]; }
    }
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
        long long p = idx / L; int c = (int)(idx - p * L) * VEC;
        Vec<VEC> g = Vec<VEC>::load(dy + (size_t)p * dy_ld + c);
        Vec<VEC> xv = Vec<VEC>::load(x + (size_t)p * x_ld + c);
        Vec<VEC> ov;
        if (act == SAUNET_ACT_RELU && out) ov = Vec<VEC>::load(out + (size_t)p * out_ld + c);
        Vec<VEC> dxv, drv;
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            const int cc = c + e;
            float gg = g.v[e];
            if (act == SAUNET_ACT_RELU) {
                float z = out ? ov.v[e] : fmaf(xv.v[e], state[cc], state[C + cc]);
                if (!(z > 0.f)) gg = 0.f;
            }
            drv.v[e] = gg;
            const float invstd = state[3 * C + cc];
            const float gs = (gamma ? gamma[cc] : 1.f) * invstd;
            if (training) {
                float xh = (xv.v[e] - state[2 * C + cc]) * invstd;
                float m1 = (float)(red[cc] * inv_n), m2 = (float)(red[C + cc] * inv_n);
                dxv.v[e] = gs * (gg - m1 - xh * m2);
            } else {
                dxv.v[e] = gs * gg;
            }
        }
        if (dx) {
            float* dp = dx + (size_t)p * dx_ld + c;
            if (dx_acc) { Vec<VEC> o = Vec<VEC>::load(dp);
#pragma unroll
                for (int e = 0; e < VEC; ++e) dxv.v[e] += o.v[e]; }
            dxv.store(dp);
        }
        if (dres) {
            float* dp = dres + (size_t)p * dres_ld + c;
            if (dres_acc) { Vec<VEC> o = Vec<VEC>::load(dp);
#pragma unroll
                for (int e = 0; e < VEC; ++e) drv.v[e] += o.v[e]; }
            drv.store(dp);
        }
    }
}


// ---- BatchNorm backward, 4-channel vector path -------------------------------------------------------------------
// Thread = one fixed 4-channel group (its BN constants live in registers) x a strided set of pixels; 4 pixels per
// iteration with all 128-bit loads issued before the math (8-12 loads in flight per thread).  Per-thread sums are
// fp32 over at most 32 pixels, then folded into fp64.
__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ld4rw(const float* p) { return *reinterpret_cast<const float4*>(p); }

struct BnGeom {
    int lanes, rows;
    __device__ BnGeom(int L) { lanes = L < 256 ? L : 256; rows = 256 / lanes; }
};

template <bool RELU, bool HAS_OUT>
__device__ __forceinline__ float4 masked_grad(float4 g, float4 xv, float4 ov, float4 sc, float4 sh) {
    if (RELU) {
        float4 z;
        if (HAS_OUT) z = ov;
        else { z.x = fmaf(xv.x, sc.x, sh.x); z.y = fmaf(xv.y, sc.y, sh.y); z.z = fmaf(xv.z, sc.z, sh.z); z.w = fmaf(xv.w, sc.w, sh.w); }
        if (!(z.x > 0.f)) g.x = 0.f;
        if (!(z.y > 0.f)) g.y = 0.f;
        if (!(z.z > 0.f)) g.z = 0.f;
        if (!(z.w > 0.f)) g.w = 0.f;
    }
    return g;
}

template <bool RELU, bool HAS_OUT>
__global__ void __launch_bounds__(256, 2) bn_bwd_reduce4_kernel(const float* __restrict__ dy, int dy_ld, const float* __restrict__ x, int x_ld,
                                                             const float* __restrict__ out, int out_ld, const float* __restrict__ state,
                                                             int C, long long npix, double* __restrict__ red) {
    __shared__ double sred[8 * 256];
    const int L = C >> 2;
    const BnGeom gm(L);
    const int tid = threadIdx.x, l = tid % gm.lanes, r = tid / gm.lanes;
    const bool active = r < gm.rows;
    const long long stride = (long long)gridDim.x * gm.rows;
    for (int l0 = 0; l0 < L; l0 += gm.lanes) {
        const int lg = l0 + l;
        const bool on = active && lg < L;
        const int c = lg * 4;
        double a1[4] = {0.0, 0.0, 0.0, 0.0}, a2[4] = {0.0, 0.0, 0.0, 0.0};
        if (on) {
            const float4 sc = ld4(state + c), sh = ld4(state + C + c), mean = ld4(state + 2 * C + c), istd = ld4(state + 3 * C + c);
            long long p = (long long)blockIdx.x * gm.rows + r;
            while (p < npix) {
                float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
                for (int it = 0; it < 8 && p < npix; ++it, p += 4 * stride) {
                    float4 g[4], xv[4], ov[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const long long pp = p + u * stride;
                        const bool ok = pp < npix;
                        g[u] = ok ? ld4(dy + (size_t)pp * dy_ld + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                        xv[u] = ok ? ld4(x + (size_t)pp * x_ld + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                        if (RELU && HAS_OUT) ov[u] = ok ? ld4(out + (size_t)pp * out_ld + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float4 gg = masked_grad<RELU, HAS_OUT>(g[u], xv[u], ov[u], sc, sh);
                        s1[0] += gg.x; s1[1] += gg.y; s1[2] += gg.z; s1[3] += gg.w;
                        s2[0] = fmaf(gg.x, (xv[u].x - mean.x) * istd.x, s2[0]); s2[1] = fmaf(gg.y, (xv[u].y - mean.y) * istd.y, s2[1]);
                        s2[2] = fmaf(gg.z, (xv[u].z - mean.z) * istd.z, s2[2]); s2[3] = fmaf(gg.w, (xv[u].w - mean.w) * istd.w, s2[3]);
                    }
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) { a1[e] += (double)s1[e]; a2[e] += (double)s2[e]; }
            }
        }
        if (gm.rows > 1) {
            __syncthreads();
#pragma unroll
            for (int e = 0; e < 4; ++e) { sred[e * 256 + tid] = a1[e]; sred[(4 + e) * 256 + tid] = a2[e]; }
            __syncthreads();
            if (r == 0 && lg < L) {
                for (int rr = 1; rr < gm.rows; ++rr)
#pragma unroll
                    for (int e = 0; e < 4; ++e) { a1[e] += sred[e * 256 + rr * gm.lanes + l]; a2[e] += sred[(4 + e) * 256 + rr * gm.lanes + l]; }
            }
        }
        if (r == 0 && lg < L) {
#pragma unroll
            for (int e = 0; e < 4; ++e) { atomicAdd(red + c + e, a1[e]); atomicAdd(red + C + c + e, a2[e]); }
        }
    }
}

template <bool RELU, bool HAS_OUT>
__global__ void __launch_bounds__(256, 2) bn_bwd_apply4_kernel(const float* __restrict__ dy, int dy_ld, const float* __restrict__ x, int x_ld,
                                                            const float* __restrict__ out, int out_ld, const float* __restrict__ state,
                                                            const float* __restrict__ gamma, const double* __restrict__ red, int C,
                                                            long long npix, int training, float* dx, int dx_ld, int dx_acc,
                                                            float* dres, int dres_ld, int dres_acc, float* dgamma, float* dbeta) {
    const int L = C >> 2;
    const BnGeom gm(L);
    const int tid = threadIdx.x, l = tid % gm.lanes, r = tid / gm.lanes;
    if (blockIdx.x == 0 && dgamma) {
        for (int c = tid; c < C; c += blockDim.x) { dbeta[c] += (float)red[c]; dgamma[c] += (float)red[C + c]; }
    }
    if (r >= gm.rows) return;
    const long long stride = (long long)gridDim.x * gm.rows;
    const double inv_n = 1.0 / (double)npix;
    for (int l0 = 0; l0 < L; l0 += gm.lanes) {
        const int lg = l0 + l;
        if (lg >= L) continue;
        const int c = lg * 4;
        const float4 sc = ld4(state + c), sh = ld4(state + C + c), mean = ld4(state + 2 * C + c), istd = ld4(state + 3 * C + c);
        // dx = gs*g - gs*m1 - (gs*m2*invstd) * (x - mean)
        float gs[4], gm1[4], kk[4];
        {
            const float is[4] = {istd.x, istd.y, istd.z, istd.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                gs[e] = (gamma ? __ldg(gamma + c + e) : 1.f) * is[e];
                const float m1 = training ? (float)(red[c + e] * inv_n) : 0.f, m2 = training ? (float)(red[C + c + e] * inv_n) : 0.f;
                gm1[e] = gs[e] * m1; kk[e] = gs[e] * m2 * is[e];
            }
        }
        for (long long p = (long long)blockIdx.x * gm.rows + r; p < npix; p += 2 * stride) {
            float4 g[2], xv[2], ov[2], od[2], orr[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const long long pp = p + u * stride;
                if (pp < npix) {
                    g[u] = ld4(dy + (size_t)pp * dy_ld + c);
                    xv[u] = ld4(x + (size_t)pp * x_ld + c);
                    if (RELU && HAS_OUT) ov[u] = ld4(out + (size_t)pp * out_ld + c);
                    if (dx && dx_acc) od[u] = ld4rw(dx + (size_t)pp * dx_ld + c);
                    if (dres && dres_acc) orr[u] = ld4rw(dres + (size_t)pp * dres_ld + c);
                }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const long long pp = p + u * stride;
                if (pp < npix) {
                    const float4 gg = masked_grad<RELU, HAS_OUT>(g[u], xv[u], ov[u], sc, sh);
                    if (dx) {
                        float4 o;
                        o.x = fmaf(gs[0], gg.x, fmaf(-kk[0], xv[u].x - mean.x, -gm1[0]));
                        o.y = fmaf(gs[1], gg.y, fmaf(-kk[1], xv[u].y - mean.y, -gm1[1]));
                        o.z = fmaf(gs[2], gg.z, fmaf(-kk[2], xv[u].z - mean.z, -gm1[2]));
                        o.w = fmaf(gs[3], gg.w, fmaf(-kk[3], xv[u].w - mean.w, -gm1[3]));
                        if (dx_acc) { o.x += od[u].x; o.y += od[u].y; o.z += od[u].z; o.w += od[u].w; }
                        *reinterpret_cast<float4*>(dx + (size_t)pp * dx_ld + c) = o;
                    }
                    if (dres) {
                        float4 o = gg;
                        if (dres_acc) { o.x += orr[u].x; o.y += orr[u].y; o.z += orr[u].z; o.w += orr[u].w; }
                        *reinterpret_cast<float4*>(dres + (size_t)pp * dres_ld + c) = o;
                    }
                }
            }
        }
    }
}


// per-channel sum / sum of squares, 4-channel vector path (same structure as bn_bwd_reduce4_kernel)
__global__ void __launch_bounds__(256, 2) channel_stats4_kernel(const float* __restrict__ x, int x_ld, int C, long long npix,
                                                                double* __restrict__ sum, double* __restrict__ sumsq) {
    __shared__ double sred[8 * 256];
    const int L = C >> 2;
    const BnGeom gm(L);
    const int tid = threadIdx.x, l = tid % gm.lanes, r = tid / gm.lanes;
    const bool active = r < gm.rows;
    const long long stride = (long long)gridDim.x * gm.rows;
    for (int l0 = 0; l0 < L; l0 += gm.lanes) {
        const int lg = l0 + l;
        const bool on = active && lg < L;
        const int c = lg * 4;
        double a1[4] = {0.0, 0.0, 0.0, 0.0}, a2[4] = {0.0, 0.0, 0.0, 0.0};
        if (on) {
            long long p = (long long)blockIdx.x * gm.rows + r;
            while (p < npix) {
                float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
                for (int it = 0; it < 8 && p < npix; ++it, p += 4 * stride) {
                    float4 v[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const long long pp = p + u * stride;
                        v[u] = pp < npix ? ld4(x + (size_t)pp * x_ld + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        s1[0] += v[u].x; s1[1] += v[u].y; s1[2] += v[u].z; s1[3] += v[u].w;
                        s2[0] = fmaf(v[u].x, v[u].x, s2[0]); s2[1] = fmaf(v[u].y, v[u].y, s2[1]);
                        s2[2] = fmaf(v[u].z, v[u].z, s2[2]); s2[3] = fmaf(v[u].w, v[u].w, s2[3]);
                    }
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) { a1[e] += (double)s1[e]; a2[e] += (double)s2[e]; }
            }
        }
        if (gm.rows > 1) {
            __syncthreads();
#pragma unroll
            for (int e = 0; e < 4; ++e) { sred[e * 256 + tid] = a1[e]; sred[(4 + e) * 256 + tid] = a2[e]; }
            __syncthreads();
            if (r == 0 && lg < L) {
                for (int rr = 1; rr < gm.rows; ++rr)
#pragma unroll
                    for (int e = 0; e < 4; ++e) { a1[e] += sred[e * 256 + rr * gm.lanes + l]; a2[e] += sred[(4 + e) * 256 + rr * gm.lanes + l]; }
            }
        }
        if (r == 0 && lg < L) {
#pragma unroll
            for (int e = 0; e < 4; ++e) { atomicAdd(sum + c + e, a1[e]); atomicAdd(sumsq + c + e, a2[e]); }
        }
    }
}

// y = act(x*scale + shift (+ residual)), 4-channel vector path: thread = fixed channel group (constants in registers)
// x strided pixels, 4 pixels per iteration with the loads issued first
__global__ void __launch_bounds__(256, 2) affine_act4_kernel(const float* __restrict__ x, int x_ld, const float* __restrict__ scale,
                                                             const float* __restrict__ shift, const float* __restrict__ res, int r_ld,
                                                             float* __restrict__ y, int y_ld, int C, long long npix, int act) {
    const int L = C >> 2;
    const BnGeom gm(L);
    const int tid = threadIdx.x, l = tid % gm.lanes, r = tid / gm.lanes;
    if (r >= gm.rows) return;
    const long long stride = (long long)gridDim.x * gm.rows;
    for (int l0 = 0; l0 < L; l0 += gm.lanes) {
        const int lg = l0 + l;
        if (lg >= L) continue;
        const int c = lg * 4;
        const float4 sc = scale ? ld4(scale + c) : make_float4(1.f, 1.f, 1.f, 1.f);
        const float4 sh = scale ? ld4(shift + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        for (long long p = (long long)blockIdx.x * gm.rows + r; p < npix; p += 4 * stride) {
            float4 v[4], rv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const long long pp = p + u * stride;
                if (pp < npix) {
                    v[u] = ld4(x + (size_t)pp * x_ld + c);
                    if (res) rv[u] = ld4(res + (size_t)pp * r_ld + c);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const long long pp = p + u * stride;
                if (pp < npix) {
                    float4 t = make_float4(fmaf(v[u].x, sc.x, sh.x), fmaf(v[u].y, sc.y, sh.y), fmaf(v[u].z, sc.z, sh.z), fmaf(v[u].w, sc.w, sh.w));
                    if (res) { t.x += rv[u].x; t.y += rv[u].y; t.z += rv[u].z; t.w += rv[u].w; }
                    t.x = apply_act(t.x, act); t.y = apply_act(t.y, act); t.z = apply_act(t.z, act); t.w = apply_act(t.w, act);
                    *reinterpret_cast<float4*>(y + (size_t)pp * y_ld + c) = t;
                }
            }
        }
    }
}

static inline int bn_blocks(int C, long long npix) {
    const int L = C / 4, lanes = L < 256 ? L : 256, rows = 256 / lanes;
    long long want = (npix + (long long)rows * 8 - 1) / ((long long)rows * 8);        // >= 8 pixels per thread when possible
    long long cap = (long long)kNumSMs * 2;          // two resident CTAs per SM (<= 128 registers): one wave
    if (want < 1) want = 1;
    return (int)(want > cap ? cap : want);
}

static inline int ew_blocks(long long n) {
    long long b = (n + 255) / 256; long long cap = (long long)kNumSMs * 16;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}
static inline bool vec_ok(int C, std::initializer_list<std::pair<const void*, int>> ts) {
    if (C % 4) return false;
    for (auto& t : ts) if (t.first && (!aligned16(t.first) || (t.second % 4))) return false;
    return true;
}

}  // namespace saunet

using namespace saunet;

extern "C" int saunet_channel_stats(const float* x, int ld, int C, long long npix, double* sum, double* sumsq, void* stream) {
    SAUNET_CHECK_ARG(x && sum && sumsq && C > 0 && npix > 0 && ld >= C, SAUNET_ERR_BAD_SHAPE, "channel_stats: bad args");
    SAUNET_CHECK_ARG(sumsq >= sum + C, SAUNET_ERR_BAD_SHAPE, "channel_stats: sumsq must follow sum by >= C doubles");
    bool vec = vec_ok(C, {{x, ld}});
    if (vec) {
        channel_stats4_kernel<<<bn_blocks(C, npix), 256, 0, (cudaStream_t)stream>>>(x, ld, C, npix, sum, sumsq);
        SAUNET_CHECK_LAUNCH("channel_stats4_kernel");
        return SAUNET_OK;
    }
    StatsOp<1> o1{x, ld}; StatsOp<4> o4{x, ld};
    return launch_reduce(o1, o4, vec, C, npix, sum, (long long)(sumsq - sum), (cudaStream_t)stream, "channel_stats_kernel");
}

__global__ void add_d2f_kernel(const double* __restrict__ src, float* dst, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] += (float)src[i];
}
extern "C" int saunet_add_d2f(const double* src, float* dst, int n, void* stream) {
    SAUNET_CHECK_ARG(src && dst && n > 0, SAUNET_ERR_BAD_SHAPE, "add_d2f: bad args");
    add_d2f_kernel<<<cdiv(n, 128), 128, 0, (cudaStream_t)stream>>>(src, dst, n);
    SAUNET_CHECK_LAUNCH("add_d2f_kernel");
    return SAUNET_OK;
}

extern "C" int saunet_bn_finalize(const double* sum, const double* sumsq, double count, const float* gamma, const float* beta,
                                  float* running_mean, float* running_var, float momentum, float eps, int training, int C,
                                  float* state, void* stream) {
    SAUNET_CHECK_ARG(C > 0 && state, SAUNET_ERR_BAD_SHAPE, "bn_finalize: bad args");
    SAUNET_CHECK_ARG(training ? (sum && sumsq && count > 0) : (running_mean && running_var), SAUNET_ERR_BAD_SHAPE,
                     "bn_finalize: missing statistics");
    bn_finalize_kernel<<<cdiv(C, 128), 128, 0, (cudaStream_t)stream>>>(sum, sumsq, count, gamma, beta, running_mean, running_var,
                                                                       momentum, eps, training, C, state);
    SAUNET_CHECK_LAUNCH("bn_finalize_kernel");
    return SAUNET_OK;
}

extern "C" int saunet_affine_act(const float* x, int x_ld, const float* scale, const float* shift, const float* residual,
                                 int r_ld, float* y, int y_ld, int C, long long npix, int act, void* stream) {
    SAUNET_CHECK_ARG(x && y && C > 0 && npix > 0 && x_ld >= C && y_ld >= C, SAUNET_ERR_BAD_SHAPE, "affine_act: bad args");
    SAUNET_CHECK_ARG((scale == nullptr) == (shift == nullptr), SAUNET_ERR_BAD_SHAPE, "affine_act: scale/shift mismatch");
    bool vec = vec_ok(C, {{x, x_ld}, {y, y_ld}, {residual, r_ld}});
    if (vec && (!scale || (aligned16(scale) && aligned16(shift)))) {
        affine_act4_kernel<<<bn_blocks(C, npix), 256, 0, (cudaStream_t)stream>>>(x, x_ld, scale, shift, residual, r_ld, y, y_ld, C, npix, act);
        SAUNET_CHECK_LAUNCH("affine_act4_kernel");
        return SAUNET_OK;
    }
    if (vec) affine_act_kernel<4><<<ew_blocks(npix * (C / 4)), 256, 0, (cudaStream_t)stream>>>(x, x_ld, scale, shift, residual, r_ld, y, y_ld, C, npix, act);
    else affine_act_kernel<1><<<ew_blocks(npix * C), 256, 0, (cudaStream_t)stream>>>(x, x_ld, scale, shift, residual, r_ld, y, y_ld, C, npix, act);
    SAUNET_CHECK_LAUNCH("affine_act_kernel");
    return SAUNET_OK;
}

extern "C" int saunet_bn_bwd_reduce(const float* dy, int dy_ld, const float* x, int x_ld, const float* out, int out_ld,
                                    const float* state, int C, long long npix, int act, double* red, void* stream) {
    SAUNET_CHECK_ARG(dy && x && state && red && C > 0 && npix > 0, SAUNET_ERR_BAD_SHAPE, "bn_bwd_reduce: bad args");
    bool vec = vec_ok(C, {{dy, dy_ld}, {x, x_ld}, {out, out_ld}}) && aligned16(state);
    if (vec && !SAUNET_ENV_FLAG("SAUNET_NO_BN4")) {
        const int blocks = bn_blocks(C, npix);
        cudaStream_t st = (cudaStream_t)stream;
        if (act != SAUNET_ACT_RELU) bn_bwd_reduce4_kernel<false, false><<<blocks, 256, 0, st>>>(dy, dy_ld, x, x_ld, out, out_ld, state, C, npix, red);
        else if (out) bn_bwd_reduce4_kernel<true, true><<<blocks, 256, 0, st>>>(dy, dy_ld, x, x_ld, out, out_ld, state, C, npix, red);
        else bn_bwd_reduce4_kernel<true, false><<<blocks, 256, 0, st>>>(dy, dy_ld, x, x_ld, out, out_ld, state, C, npix, red);
        SAUNET_CHECK_LAUNCH("bn_bwd_reduce4_kernel");
        return SAUNET_OK;
    }
    BnBwdOp<1> o1{dy, dy_ld, x, x_ld, out, out_ld, state, C, act};
    BnBwdOp<4> o4{dy, dy_ld, x, x_ld, out, out_ld, state, C, act};
    return launch_reduce(o1, o4, vec, C, npix, red, (long long)C, (cudaStream_t)stream, "bn_bwd_reduce_kernel");
}

namespace saunet {
__global__ void bn_fused_finish_kernel(const double* __restrict__ sums, const float* __restrict__ state, double inv_n, int training,
                                       int C, float* dgamma, float* dbeta, double* ab, int ab_ld) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double s1 = sums[c], s2 = sums[C + c];
    const double scale = state[c], mean = state[2 * C + c], istd = state[3 * C + c];
    if (dbeta) dbeta[c] += (float)s1;
    if (dgamma) dgamma[c] += (float)(istd * s2);
    if (training && ab) {
        const double B = scale * istd * istd * s2 * inv_n;
        ab[c] += scale * s1 * inv_n - B * mean;
        ab[ab_ld + c] += B;
    }
}

// thread = 4 channels x strided pixels (C % 4 == 0, 16-byte aligned rows) or 1 channel
template <int V>
__global__ void __launch_bounds__(256) bn_fixup_kernel(float* __restrict__ g, int g_ld, const float* __restrict__ x, int x_ld,
                                                       const double* __restrict__ ab, int ab_ld, int C, long long npix) {
    const int L = C / V;
    const long long total = npix * L;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / L; const int c = (int)(i - p * L) * V;
        if (V == 4) {
            float4 gv = *reinterpret_cast<const float4*>(g + (size_t)p * g_ld + c);
            const float4 xv = __ldg(reinterpret_cast<const float4*>(x + (size_t)p * x_ld + c));
            gv.x -= (float)ab[c] + (float)ab[ab_ld + c] * xv.x;
            gv.y -= (float)ab[c + 1] + (float)ab[ab_ld + c + 1] * xv.y;
            gv.z -= (float)ab[c + 2] + (float)ab[ab_ld + c + 2] * xv.z;
            gv.w -= (float)ab[c + 3] + (float)ab[ab_ld + c + 3] * xv.w;
            *reinterpret_cast<float4*>(g + (size_t)p * g_ld + c) = gv;
        } else {
            g[(size_t)p * g_ld + c] -= (float)ab[c] + (float)ab[ab_ld + c] * __ldg(x + (size_t)p * x_ld + c);
        }
    }
}
}  // namespace saunet

extern "C" int saunet_bn_fused_finish(const double* sums, const float* state, double count, int training, int C, float* dgamma,
                                      float* dbeta, double* ab, int ab_ld, void* stream) {
    SAUNET_CHECK_ARG(sums && state && C > 0 && count > 0, SAUNET_ERR_BAD_SHAPE, "bn_fused_finish: bad args");
    SAUNET_CHECK_ARG(!training || (ab && ab_ld >= C), SAUNET_ERR_BAD_SHAPE, "bn_fused_finish: training mode needs ab[2][>=C]");
    bn_fused_finish_kernel<<<cdiv(C, 128), 128, 0, (cudaStream_t)stream>>>(sums, state, 1.0 / count, training, C, dgamma, dbeta, ab, ab_ld);
    SAUNET_CHECK_LAUNCH("bn_fused_finish_kernel");
    return SAUNET_OK;
}

extern "C" int saunet_bn_fixup(float* g, int g_ld, const float* x, int x_ld, const double* ab, int ab_ld, int C, long long npix,
                               void* stream) {
    SAUNET_CHECK_ARG(g && x && ab && C > 0 && npix > 0 && g_ld >= C && x_ld >= C, SAUNET_ERR_BAD_SHAPE, "bn_fixup: bad args");
    const bool vec = C % 4 == 0 && g_ld % 4 == 0 && x_ld % 4 == 0 && aligned16(g) && aligned16(x);
    if (vec) bn_fixup_kernel<4><<<ew_blocks(npix * (C / 4)), 256, 0, (cudaStream_t)stream>>>(g, g_ld, x, x_ld, ab, ab_ld, C, npix);
    else bn_fixup_kernel<1><<<ew_blocks(npix * C), 256, 0, (cudaStream_t)stream>>>(g, g_ld, x, x_ld, ab, ab_ld, C, npix);
    SAUNET_CHECK_LAUNCH("bn_fixup_kernel");
    return SAUNET_OK;
}

extern "C" int saunet_bn_bwd_apply(const float* dy, int dy_ld, const float* x, int x_ld, const float* out, int out_ld,
                                   const float* state, const float* gamma, const double* red, int C, long long npix, int act,
                                   int training, float* dx, int dx_ld, int dx_acc, float* dres, int dres_ld, int dres_acc,
                                   float* dgamma, float* dbeta, void* stream) {
    SAUNET_CHECK_ARG(dy && x && state && red && C > 0 && npix > 0, SAUNET_ERR_BAD_SHAPE, "bn_bwd_apply: bad args");
    SAUNET_CHECK_ARG((dgamma == nullptr) == (dbeta == nullptr), SAUNET_ERR_BAD_SHAPE, "bn_bwd_apply: dgamma/dbeta mismatch");
    bool vec = vec_ok(C, {{dy, dy_ld}, {x, x_ld}, {out, out_ld}, {dx, dx_ld}, {dres, dres_ld}});
    if (vec && aligned16(state) && !SAUNET_ENV_FLAG("SAUNET_NO_BN4")) {
        const int blocks = bn_blocks(C, npix);
        cudaStream_t st = (cudaStream_t)stream;
#define SAUNET_BN_APPLY(R, O) bn_bwd_apply4_kernel<R, O><<<blocks, 256, 0, st>>>(dy, dy_ld, x, x_ld, out, out_ld, state, gamma, red, C, npix, training, dx, dx_ld, dx_acc, dres, dres_ld, dres_acc, dgamma, dbeta)
        if (act != SAUNET_ACT_RELU) SAUNET_BN_APPLY(false, false);
        else if (out) SAUNET_BN_APPLY(true, true);
        else SAUNET_BN_APPLY(true, false);
#undef SAUNET_BN_APPLY
        SAUNET_CHECK_LAUNCH("bn_bwd_apply4_kernel");
        return SAUNET_OK;
    }
    if (vec) bn_bwd_apply_kernel<4><<<ew_blocks(npix * (C / 4)), 256, 0, (cudaStream_t)stream>>>(dy, dy_ld, x, x_ld, out, out_ld, state, gamma, red, C, npix, act, training, dx, dx_ld, dx_acc, dres, dres_ld, dres_acc, dgamma, dbeta);
    else bn_bwd_apply_kernel<1><<<ew_blocks(npix * C), 256, 0, (cudaStream_t)stream>>>(dy, dy_ld, x, x_ld, out, out_ld, state, gamma, red, C, npix, act, training, dx, dx_ld, dx_acc, dres, dres_ld, dres_acc, dgamma, dbeta);
    SAUNET_CHECK_LAUNCH("bn_bwd_apply_kernel");
    return SAUNET_OK;
}
