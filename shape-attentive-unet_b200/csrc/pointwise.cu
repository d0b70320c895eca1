// pointwise.cu -- HBM-bound resampling / gating / layout kernels on NHWC fp32.
#include "common.cuh"

namespace saunet {

static inline int ew_blocks(long long n) {
    long long b = (n + 255) / 256; long long cap = (long long)kNumSMs * 16;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}
static inline bool vec4(int C, const void* a, int lda, const void* b = nullptr, int ldb = 0, const void* c = nullptr, int ldc = 0) {
    if (C % 4) return false;
    if (a && (!aligned16(a) || lda % 4)) return false;
    if (b && (!aligned16(b) || ldb % 4)) return false;
    if (c && (!aligned16(c) || ldc % 4)) return false;
    return true;
}

template <int VEC> struct V;
template <> struct V<1> {
    float v[1];
    __device__ static V load(const float* p) { V r; r.v[0] = __ldg(p); return r; }
    __device__ static V loadrw(const float* p) { V r; r.v[0] = *p; return r; }
    __device__ void store(float* p) const { p[0] = v[0]; }
};
template <> struct V<4> {
    float v[4];
    __device__ static V load(const float* p) { float4 t = __ldg(reinterpret_cast<const float4*>(p)); V r; r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w; return r; }
    __device__ static V loadrw(const float* p) { float4 t = *reinterpret_cast<const float4*>(p); V r; r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w; return r; }
    __device__ void store(float* p) const { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
};

// ---- bilinear, align_corners=True -----------------------------------------------------------------
__device__ __forceinline__ void src_index(float scale, int o, int in, int& i0, int& i1, float& l0, float& l1) {
    float s = scale * (float)o;
    i0 = (int)s; if (i0 > in - 1) i0 = in - 1;
    i1 = i0 + (i0 < in - 1 ? 1 : 0);
    l1 = s - (float)i0; l0 = 1.f - l1;
}

template <int VEC>
__global__ void __launch_bounds__(256) bilinear_fwd_kernel(const float* __restrict__ x, int x_ld, int B, int Hin, int Win, int C,
                                                           float* __restrict__ y, int y_ld, int Hout, int Wout, float sh, float sw) {
    const int L = C / VEC;
    const long long n = (long long)B * Hout * Wout * L;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
        int c = (int)(idx % L) * VEC; long long p = idx / L;
        int ox = (int)(p % Wout); long long q = p / Wout; int oy = (int)(q % Hout); int b = (int)(q / Hout);
        int y0, y1, x0, x1; float ly0, ly1, lx0, lx1;
        src_index(sh, oy, Hin, y0, y1, ly0, ly1);
        src_index(sw, ox, Win, x0, x1, lx0, lx1);
        const float* base = x + (size_t)b * Hin * Win * x_ld + c;
        V<VEC> v00 = V<VEC>::load(base + ((size_t)y0 * Win + x0) * x_ld);
        V<VEC> v01 = V<VEC>::load(base + ((size_t)y0 * Win + x1) * x_ld);
        V<VEC> v10 = V<VEC>::load(base + ((size_t)y1 * Win + x0) * x_ld);
        V<VEC> v11 = V<VEC>::load(base + ((size_t)y1 * Win + x1) * x_ld);
        V<VEC> o;
#pragma unroll
        for (int e = 0; e < VEC; ++e)
            o.v[e] = ly0 * (lx0 * v00.v[e] + lx1 * v01.v[e]) + ly1 * (lx0 * v10.v[e] + lx1 * v11.v[e]);
        o.store(y + (size_t)p * y_ld + c);
    }
}

// gather form of the backward: every input pixel sums the output pixels that read it (deterministic,
// no atomics, no memset).
template <int VEC>
__global__ void __launch_bounds__(256) bilinear_bwd_kernel(const float* __restrict__ dy, int dy_ld, int B, int Hin, int Win, int C,
                                                           float* dx, int dx_ld, int Hout, int Wout, float sh, float sw, int accumulate) {
    const int L = C / VEC;
    const long long n = (long long)B * Hin * Win * L;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
        int c = (int)(idx % L) * VEC; long long p = idx / L;
        int ix = (int)(p % Win); long long q = p / Win; int iy = (int)(q % Hin); int b = (int)(q / Hin);
        int oylo = 0, oyhi = Hout - 1, oxlo = 0, oxhi = Wout - 1;
        if (sh > 0.f) { oylo = max(0, (int)floorf((float)(iy - 1) / sh) - 1); oyhi = min(Hout - 1, (int)ceilf((float)(iy + 1) / sh) + 1); }
        if (sw > 0.f) { oxlo = max(0, (int)floorf((float)(ix - 1) / sw) - 1); oxhi = min(Wout - 1, (int)ceilf((float)(ix + 1) / sw) + 1); }
        float acc[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[e] = 0.f;
        const float* base = dy + (size_t)b * Hout * Wout * dy_ld + c;
        for (int oy = oylo; oy <= oyhi; ++oy) {
            int y0, y1; float ly0, ly1;
            src_index(sh, oy, Hin, y0, y1, ly0, ly1);
            float wy = (y0 == iy ? ly0 : 0.f) + (y1 == iy ? ly1 : 0.f);
            if (wy == 0.f) continue;
            for (int ox = oxlo; ox <= oxhi; ++ox) {
                int x0, x1; float lx0, lx1;
                src_index(sw, ox, Win, x0, x1, lx0, lx1);
                float wx = (x0 == ix ? lx0 : 0.f) + (x1 == ix ? lx1 : 0.f);
                if (wx == 0.f) continue;
                V<VEC> g = V<VEC>::load(base + ((size_t)oy * Wout + ox) * dy_ld);
                float w = wy * wx;
#pragma unroll
                for (int e = 0; e < VEC; ++e) acc[e] = fmaf(w, g.v[e], acc[e]);
            }
        }
        float* dp = dx + (size_t)p * dx_ld + c;
        V<VEC> o;
        if (accumulate) { o = V<VEC>::loadrw(dp);
#pragma unroll
            for (int e = 0; e < VEC; ++e) o.v[e] += acc[e]; }
        else {
#pragma unroll
            for (int e = 0; e < VEC; ++e) o.v[e] = acc[e]; }
        o.store(dp);
    }
}

// single-channel maps upsampled by a large factor (the c3/c4/c5 heads, x8 / x16): a thread-per-input-pixel gather walks
// up to 34x34 candidates serially on only B*Hin*Win threads; here one WARP owns an input pixel and its lanes split the
// candidate window, then shuffle-reduce (still deterministic, no atomics).
__global__ void __launch_bounds__(256) bilinear_bwd_c1_warp_kernel(const float* __restrict__ dy, int dy_ld, int B, int Hin, int Win,
                                                                   float* dx, int dx_ld, int Hout, int Wout, float sh, float sw, int accumulate) {
    const int lane = threadIdx.x & 31;
    const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long n = (long long)B * Hin * Win;
    for (long long p = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5; p < n; p += nw) {
        const int ix = (int)(p % Win); const long long q = p / Win; const int iy = (int)(q % Hin); const int b = (int)(q / Hin);
        int oylo = 0, oyhi = Hout - 1, oxlo = 0, oxhi = Wout - 1;
        if (sh > 0.f) { oylo = max(0, (int)floorf((float)(iy - 1) / sh) - 1); oyhi = min(Hout - 1, (int)ceilf((float)(iy + 1) / sh) + 1); }
        if (sw > 0.f) { oxlo = max(0, (int)floorf((float)(ix - 1) / sw) - 1); oxhi = min(Wout - 1, (int)ceilf((float)(ix + 1) / sw) + 1); }
        const int ww = oxhi - oxlo + 1, cnt = (oyhi - oylo + 1) * ww;
        const float* base = dy + (size_t)b * Hout * Wout * dy_ld;
        float acc = 0.f;
        for (int k = lane; k < cnt; k += 32) {
            const int oy = oylo + k / ww, ox = oxlo + k % ww;
            int y0, y1, x0, x1; float ly0, ly1, lx0, lx1;
            src_index(sh, oy, Hin, y0, y1, ly0, ly1);
            src_index(sw, ox, Win, x0, x1, lx0, lx1);
            const float w = ((y0 == iy ? ly0 : 0.f) + (y1 == iy ? ly1 : 0.f)) * ((x0 == ix ? lx0 : 0.f) + (x1 == ix ? lx1 : 0.f));
            if (w != 0.f) acc = fmaf(w, __ldg(base + ((size_t)oy * Wout + ox) * dy_ld), acc);
        }
        acc = warp_sum(acc);
        if (lane == 0) { float* dp = dx + (size_t)p * dx_ld; *dp = accumulate ? *dp + acc : acc; }
    }
}

// ---- 2x2 pools ----------------------------------------------------------------------------------------
template <int VEC, bool MAXP>
__global__ void __launch_bounds__(256) pool2_fwd_kernel(const float* __restrict__ x, int x_ld, int B, int Hin, int Win, int C,
                                                        float* __restrict__ y, int y_ld) {
    const int L = C / VEC, Ho = Hin / 2, Wo = Win / 2;
    const long long n = (long long)B * Ho * Wo * L;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
        int c = (int)(idx % L) * VEC; long long p = idx / L;
        int ox = (int)(p % Wo); long long q = p / Wo; int oy = (int)(q % Ho); int b = (int)(q / Ho);
        const float* base = x + (((size_t)b * Hin + 2 * oy) * Win + 2 * ox) * x_ld + c;
        V<VEC> a = V<VEC>::load(base), bb = V<VEC>::load(base + x_ld);
        V<VEC> cc = V<VEC>::load(base + (size_t)Win * x_ld), dd = V<VEC>::load(base + (size_t)Win * x_ld + x_ld);
        V<VEC> o;
#pragma unroll
        for (int e = 0; e < VEC; ++e)
            o.v[e] = MAXP ? fmaxf(fmaxf(a.v[e], bb.v[e]), fmaxf(cc.v[e], dd.v[e])) : ((a.v[e] + bb.v[e]) + (cc.v[e] + dd.v[e])) * 0.25f;
        o.store(y + (size_t)p * y_ld + c);
    }
}
template <int VEC, bool MAXP>
__global__ void __launch_bounds__(256) pool2_bwd_kernel(const float* __restrict__ dy, int dy_ld, const float* __restrict__ x, int x_ld,
                                                        int B, int Hin, int Win, int C, float* dx, int dx_ld, int accumulate) {
    // one thread per INPUT pixel group so odd trailing rows/cols receive an explicit zero
    const int L = C / VEC, Ho = Hin / 2, Wo = Win / 2;
    const long long n = (long long)B * Hin * Win * L;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
        int c = (int)(idx % L) * VEC; long long p = idx / L;
        int ix = (int)(p % Win); long long q = p / Win; int iy = (int)(q % Hin); int b = (int)(q / Hin);
        int oy = iy >> 1, ox = ix >> 1;
        V<VEC> o;
#pragma unroll
        for (int e = 0; e < VEC; ++e) o.v[e] = 0.f;
        if (oy < Ho && ox < Wo) {
            V<VEC> g = V<VEC>::load(dy + (((size_t)b * Ho + oy) * Wo + ox) * dy_ld + c);
            if (MAXP) {
                const float* base = x + (((size_t)b * Hin + 2 * oy) * Win + 2 * ox) * x_ld + c;
                V<VEC> w[4] = {V<VEC>::load(base), V<VEC>::load(base + x_ld), V<VEC>::load(base + (size_t)Win * x_ld),
                               V<VEC>::load(base + (size_t)Win * x_ld + x_ld)};
                const int me = (iy & 1) * 2 + (ix & 1);
#pragma unroll
                for (int e = 0; e < VEC; ++e) {
                    int am = 0; float m = w[0].v[e];
#pragma unroll
                    for (int k = 1; k < 4; ++k) if (w[k].v[e] > m) { m = w[k].v[e]; am = k; }
                    o.v[e] = (am == me) ? g.v[e] : 0.f;
                }
            } else {
#pragma unroll
                for (int e = 0; e < VEC; ++e) o.v[e] = 0.25f * g.v[e];
            }
        }
        float* dp = dx + (size_t)p * dx_ld + c;
        if (accumulate) { V<VEC> t = V<VEC>::loadrw(dp);
#pragma unroll
            for (int e = 0; e < VEC; ++e) o.v[e] += t.v[e]; }
        o.store(dp);
    }
}

// ---- global average pool --------------------------------------------------------------------------------
// grid (C/32, B, splits); block 256 = 8 pixel rows x 32 channel lanes; y pre-zeroed, float atomics.
__global__ void __launch_bounds__(256) gap_fwd_kernel(const float* __restrict__ x, int x_ld, long long HW, int C, float* y, float inv) {
    __shared__ float red[8][33];
    const int lane = threadIdx.x & 31, r = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane, b = blockIdx.y;
    const long long per = (HW + gridDim.z - 1) / gridDim.z;
    const long long p0 = blockIdx.z * per; long long p1 = p0 + per; if (p1 > HW) p1 = HW;
    float acc = 0.f;
    if (c < C) {
        const float* base = x + (size_t)b * HW * x_ld + c;
        for (long long p = p0 + r; p < p1; p += 8) acc += __ldg(base + (size_t)p * x_ld);
    }
    red[r][lane] = acc;
    __syncthreads();
    if (r == 0 && c < C) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) s += red[k][lane];
        atomicAdd(y + (size_t)b * C + c, s * inv);
    }
}
// C % 4 == 0, 16-byte aligned rows: lane = 4 channels (one 512-byte request per warp and pixel row), 4 rows in flight per
// thread.  (The scalar kernel above reads 128 bytes per request with one load in flight: 2.2 TB/s at C = 256.)
__global__ void __launch_bounds__(256) gap_fwd4_kernel(const float* __restrict__ x, int x_ld, long long HW, int C, float* y, float inv) {
    __shared__ float4 red[8][32];
    const int lane = threadIdx.x & 31, r = threadIdx.x >> 5;
    const int c = (blockIdx.x * 32 + lane) * 4, b = blockIdx.y;
    const long long per = (HW + gridDim.z - 1) / gridDim.z;
    const long long p0 = blockIdx.z * per; long long p1 = p0 + per; if (p1 > HW) p1 = HW;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < C) {
        const float* base = x + (size_t)b * HW * x_ld + c;
        long long p = p0 + r;
        for (; p + 24 < p1; p += 32) {
            const float4 a0 = __ldg(reinterpret_cast<const float4*>(base + (size_t)p * x_ld));
            const float4 a1 = __ldg(reinterpret_cast<const float4*>(base + (size_t)(p + 8) * x_ld));
            const float4 a2 = __ldg(reinterpret_cast<const float4*>(base + (size_t)(p + 16) * x_ld));
            const float4 a3 = __ldg(reinterpret_cast<const float4*>(base + (size_t)(p + 24) * x_ld));
            acc.x += (a0.x + a1.x) + (a2.x + a3.x); acc.y += (a0.y + a1.y) + (a2.y + a3.y);
            acc.z += (a0.z + a1.z) + (a2.z + a3.z); acc.w += (a0.w + a1.w) + (a2.w + a3.w);
        }
        for (; p < p1; p += 8) {
            const float4 a0 = __ldg(reinterpret_cast<const float4*>(base + (size_t)p * x_ld));
            acc.x += a0.x; acc.y += a0.y; acc.z += a0.z; acc.w += a0.w;
        }
    }
    red[r][lane] = acc;
    __syncthreads();
    if (r == 0 && c < C) {
        float4 s = red[0][lane];
#pragma unroll
        for (int k = 1; k < 8; ++k) { s.x += red[k][lane].x; s.y += red[k][lane].y; s.z += red[k][lane].z; s.w += red[k][lane].w; }
        float* yp = y + (size_t)b * C + c;
        atomicAdd(yp, s.x * inv); atomicAdd(yp + 1, s.y * inv); atomicAdd(yp + 2, s.z * inv); atomicAdd(yp + 3, s.w * inv);
    }
}
template <int VEC>
__global__ void __launch_bounds__(256) gap_bwd_kernel(const float* __restrict__ dy, int B, long long HW, int C, float* dx, int dx_ld,
                                                      int accumulate, float inv) {
    const int L = C / VEC;
    const long long n = (long long)B * HW * L;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
        int c = (int)(idx % L) * VEC; long long p = idx / L; int b = (int)(p / HW);
        V<VEC> g = V<VEC>::load(dy + (size_t)b * C + c);
        float* dp = dx + (size_t)p * dx_ld + c;
        V<VEC> o;
        if (accumulate) o = V<VEC>::loadrw(dp);
#pragma unroll
        for (int e = 0; e < VEC; ++e) o.v[e] = (accumulate ? o.v[e] : 0.f) + g.v[e] * inv;
        o.store(dp);
    }
}

template <int VEC>
__global__ void __launch_bounds__(256) act_bwd_kernel(const float* __restrict__ dy, int dy_ld, const float* __restrict__ y, int y_ld,
                                                      float* dz, int dz_ld, int C, long long npix, int act) {
    const int L = C / VEC;
    const long long n = npix * L;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
        long long p = idx / L; int c = (int)(idx - p * L) * VEC;
        V<VEC> g = V<VEC>::load(dy + (size_t)p * dy_ld + c), yv = V<VEC>::load(y + (size_t)p * y_ld + c);
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            if (act == SAUNET_ACT_SIGMOID) g.v[e] = g.v[e] * yv.v[e] * (1.f - yv.v[e]);
            else if (act == SAUNET_ACT_RELU) g.v[e] = yv.v[e] > 0.f ? g.v[e] : 0.f;
        }
        g.store(dz + (size_t)p * dz_ld + c);
    }
}

// ---- DualAttBlock combine -----------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(256) dualatt_fwd_kernel(const float* __restrict__ f, int f_ld, const float* __restrict__ S,
                                                          const float* __restrict__ cv, long long HW, int C, long long npix,
                                                          float* __restrict__ out, int o_ld) {
    const int L = C / VEC;
    const long long n = npix * L;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
        long long p = idx / L; int c = (int)(idx - p * L) * VEC; int b = (int)(p / HW);
        V<VEC> v = V<VEC>::load(f + (size_t)p * f_ld + c), cc = V<VEC>::load(cv + (size_t)b * C + c);
        const float s1 = __ldg(S + p) + 1.f;
#pragma unroll
        for (int e = 0; e < VEC; ++e) v.v[e] = s1 * (v.v[e] * cc.v[e]);
        v.store(out + (size_t)p * o_ld + c);
    }
}
// grid (chunks, B); block 256 = 8 warps; each warp walks pixels of image b, lanes walk channels.
template <int MAXCH>
__global__ void __launch_bounds__(256) dualatt_bwd_kernel(const float* __restrict__ dout, int do_ld, const float* __restrict__ f, int f_ld,
                                                          const float* __restrict__ S, const float* __restrict__ cv, long long HW, int C,
                                                          float* df, int df_ld, int df_acc, float* __restrict__ dS, float* dcv) {
    __shared__ float red[8][32 * MAXCH + 1];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, b = blockIdx.y;
    const long long per = (HW + gridDim.x - 1) / gridDim.x;
    const long long p0 = blockIdx.x * per; long long p1 = p0 + per; if (p1 > HW) p1 = HW;
    float accc[MAXCH], cvr[MAXCH];
#pragma unroll
    for (int k = 0; k < MAXCH; ++k) { accc[k] = 0.f; int c = k * 32 + lane; cvr[k] = c < C ? cv[(size_t)b * C + c] : 0.f; }
    for (long long pl = p0 + w; pl < p1; pl += 8) {
        const long long p = (long long)b * HW + pl;
        const float s1 = __ldg(S + p) + 1.f;
        float accS = 0.f;
#pragma unroll
        for (int k = 0; k < MAXCH; ++k) {
            int c = k * 32 + lane;
            if (c < C) {
                float g = __ldg(dout + (size_t)p * do_ld + c), fv = __ldg(f + (size_t)p * f_ld + c);
                float t = g * fv;
                accS = fmaf(t, cvr[k], accS);
                accc[k] = fmaf(t, s1, accc[k]);
                float d = g * s1 * cvr[k];
                float* dp = df + (size_t)p * df_ld + c;
                *dp = df_acc ? *dp + d : d;
            }
        }
        accS = warp_sum(accS);
        if (lane == 0) dS[p] = accS;
    }
#pragma unroll
    for (int k = 0; k < MAXCH; ++k) red[w][k * 32 + lane] = accc[k];
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * MAXCH; i += 256) {
        if (i < C) {
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) s += red[k][i];
            atomicAdd(dcv + (size_t)b * C + i, s);
        }
    }
}

// C % 4 == 0, 16-byte aligned rows: lane = 4 channels, all loads of a pixel issued before its stores, two pixels in flight
// per warp.  (The scalar kernel above interleaves 4-byte loads and stores: 0.6 ms = 2.7 TB/s at C = 256, 524288 pixels.)
template <int K4>
__global__ void __launch_bounds__(256) dualatt_bwd4_kernel(const float* __restrict__ dout, int do_ld, const float* __restrict__ f, int f_ld,
                                                           const float* __restrict__ S, const float* __restrict__ cv, long long HW, int C,
                                                           float* df, int df_ld, int df_acc, float* __restrict__ dS, float* dcv) {
    __shared__ float red[8][128 * K4 + 4];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, b = blockIdx.y;
    const long long per = (HW + gridDim.x - 1) / gridDim.x;
    const long long p0 = blockIdx.x * per; long long p1 = p0 + per; if (p1 > HW) p1 = HW;
    float4 accc[K4], cvr[K4];
#pragma unroll
    for (int k = 0; k < K4; ++k) {
        accc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        const int c = (k * 32 + lane) * 4;
        cvr[k] = c < C ? __ldg(reinterpret_cast<const float4*>(cv + (size_t)b * C + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (long long pl = p0 + w; pl < p1; pl += 16) {
        float4 g[2][K4], fv[2][K4], od[2][K4];
        float s1[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const long long pp = pl + 8 * u;
            const long long p = (long long)b * HW + pp;
            s1[u] = pp < p1 ? __ldg(S + p) + 1.f : 0.f;
#pragma unroll
            for (int k = 0; k < K4; ++k) {
                const int c = (k * 32 + lane) * 4;
                if (pp < p1 && c < C) {
                    g[u][k] = __ldg(reinterpret_cast<const float4*>(dout + (size_t)p * do_ld + c));
                    fv[u][k] = __ldg(reinterpret_cast<const float4*>(f + (size_t)p * f_ld + c));
                    if (df_acc) od[u][k] = *reinterpret_cast<const float4*>(df + (size_t)p * df_ld + c);
                } else {
                    g[u][k] = fv[u][k] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const long long pp = pl + 8 * u;
            if (pp >= p1) continue;                      // warp-uniform
            const long long p = (long long)b * HW + pp;
            float accS = 0.f;
#pragma unroll
            for (int k = 0; k < K4; ++k) {
                const int c = (k * 32 + lane) * 4;
                const float4 t = make_float4(g[u][k].x * fv[u][k].x, g[u][k].y * fv[u][k].y, g[u][k].z * fv[u][k].z, g[u][k].w * fv[u][k].w);
                accS += t.x * cvr[k].x + t.y * cvr[k].y + t.z * cvr[k].z + t.w * cvr[k].w;
                accc[k].x = fmaf(t.x, s1[u], accc[k].x); accc[k].y = fmaf(t.y, s1[u], accc[k].y);
                accc[k].z = fmaf(t.z, s1[u], accc[k].z); accc[k].w = fmaf(t.w, s1[u], accc[k].w);
                if (c < C) {
                    float4 d = make_float4(g[u][k].x * s1[u] * cvr[k].x, g[u][k].y * s1[u] * cvr[k].y, g[u][k].z * s1[u] * cvr[k].z, g[u][k].w * s1[u] * cvr[k].w);
                    if (df_acc) { d.x += od[u][k].x; d.y += od[u][k].y; d.z += od[u][k].z; d.w += od[u][k].w; }
                    *reinterpret_cast<float4*>(df + (size_t)p * df_ld + c) = d;
                }
            }
            accS = warp_sum(accS);
            if (lane == 0) dS[p] = accS;
        }
    }
#pragma unroll
    for (int k = 0; k < K4; ++k) {
        float* r = &red[w][(k * 32 + lane) * 4];
        r[0] = accc[k].x; r[1] = accc[k].y; r[2] = accc[k].z; r[3] = accc[k].w;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 128 * K4; i += 256) {
        if (i < C) {
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) s += red[k][i];
            atomicAdd(dcv + (size_t)b * C + i, s);
        }
    }
}

// ---- GSConv gate backward --------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rowscale_bwd_kernel(const float* __restrict__ dout, int do_ld, const float* __restrict__ out, int o_ld,
                                                           const float* __restrict__ alpha, int C, long long npix, float* __restrict__ du,
                                                           int du_ld, float* dalpha, int da_acc) {
    // one warp per pixel, lanes over channels
    const int lane = threadIdx.x & 31;
    const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long p = warp; p < npix; p += nw) {
        const float a1 = __ldg(alpha + p) + 1.f;
        const float inv = 1.f / a1;
        float acc = 0.f;
        for (int c = lane; c < C; c += 32) {
            float g = __ldg(dout + (size_t)p * do_ld + c), o = __ldg(out + (size_t)p * o_ld + c);
            du[(size_t)p * du_ld + c] = g * a1;
            acc = fmaf(g, o * inv, acc);
        }
        acc = warp_sum(acc);
        if (lane == 0) dalpha[p] = da_acc ? dalpha[p] + acc : acc;
    }
}

// vector variant: LPP = C/4 lanes per pixel (1, 2, 4 or 8), each lane one 128-bit chunk; 32/LPP pixels per warp
template <int LPP>
__global__ void __launch_bounds__(256) rowscale_bwd4_kernel(const float* __restrict__ dout, int do_ld, const float* __restrict__ out, int o_ld,
                                                            const float* __restrict__ alpha, long long npix, float* __restrict__ du,
                                                            int du_ld, float* dalpha, int da_acc) {
    const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long nthr = (long long)gridDim.x * blockDim.x;
    const int sub = (int)(tid % LPP);
    for (long long p = tid / LPP; p < npix; p += nthr / LPP) {            // (npix tail: whole LPP groups stay together)
        const float a1 = __ldg(alpha + p) + 1.f;
        const float inv = 1.f / a1;
        const float4 g = __ldg(reinterpret_cast<const float4*>(dout + (size_t)p * do_ld) + sub);
        const float4 o = __ldg(reinterpret_cast<const float4*>(out + (size_t)p * o_ld) + sub);
        *(reinterpret_cast<float4*>(du + (size_t)p * du_ld) + sub) = make_float4(g.x * a1, g.y * a1, g.z * a1, g.w * a1);
        float acc = (g.x * o.x + g.y * o.y + g.z * o.z + g.w * o.w) * inv;
#pragma unroll
        for (int off = LPP / 2; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (sub == 0) dalpha[p] = da_acc ? dalpha[p] + acc : acc;
    }
}

template <int VEC>
__global__ void __launch_bounds__(256) copy_slice_kernel(const float* __restrict__ src, int s_ld, float* dst, int d_ld, int C, long long npix, int accumulate) {
    const int L = C / VEC;
    const long long n = npix * L;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
        long long p = idx / L; int c = (int)(idx - p * L) * VEC;
        V<VEC> v = V<VEC>::load(src + (size_t)p * s_ld + c);
        float* dp = dst + (size_t)p * d_ld + c;
        if (accumulate) { V<VEC> o = V<VEC>::loadrw(dp);
#pragma unroll
            for (int e = 0; e < VEC; ++e) v.v[e] += o.v[e]; }
        v.store(dp);
    }
}

// ---- layout conversion: 32x32 smem tile transpose per image -----------------------------------------------
// TO_NHWC: src [B][C][HW] -> dst [B][HW][ld];  else src [B][HW][ld] -> dst [B][C][HW]
template <bool TO_NHWC>
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int ld, int C, long long HW) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const long long p0 = (long long)blockIdx.x * 32; const int c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    if (TO_NHWC) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int c = c0 + ty + k * 8; long long p = p0 + tx;
            tile[ty + k * 8][tx] = (c < C && p < HW) ? __ldg(src + ((size_t)b * C + c) * HW + p) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            long long p = p0 + ty + k * 8; int c = c0 + tx;
            if (c < C && p < HW) dst[((size_t)b * HW + p) * ld + c] = tile[tx][ty + k * 8];
        }
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            long long p = p0 + ty + k * 8; int c = c0 + tx;
            tile[ty + k * 8][tx] = (c < C && p < HW) ? __ldg(src + ((size_t)b * HW + p) * ld + c) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int c = c0 + ty + k * 8; long long p = p0 + tx;
            if (c < C && p < HW) dst[((size_t)b * C + c) * HW + p] = tile[tx][ty + k * 8];
        }
    }
}

}  // namespace saunet

using namespace saunet;
#define ST ((cudaStream_t)stream)

extern "C" int saunet_bilinear_fwd(const float* x, int x_ld, int B, int Hin, int Win, int C, float* y, int y_ld, int Hout, int Wout, void* stream) {
    SAUNET_CHECK_ARG(x && y && B > 0 && Hin > 0 && Win > 0 && C > 0 && Hout > 0 && Wout > 0, SAUNET_ERR_BAD_SHAPE, "bilinear_fwd: bad args");
    float sh = Hout > 1 ? (float)(Hin - 1) / (float)(Hout - 1) : 0.f, sw = Wout > 1 ? (float)(Win - 1) / (float)(Wout - 1) : 0.f;
    long long n = (long long)B * Hout * Wout;
    if (vec4(C, x, x_ld, y, y_ld)) bilinear_fwd_kernel<4><<<ew_blocks(n * (C / 4)), 256, 0, ST>>>(x, x_ld, B, Hin, Win, C, y, y_ld, Hout, Wout, sh, sw);
    else bilinear_fwd_kernel<1><<<ew_blocks(n * C), 256, 0, ST>>>(x, x_ld, B, Hin, Win, C, y, y_ld, Hout, Wout, sh, sw);
    SAUNET_CHECK_LAUNCH("bilinear_fwd_kernel");
    return SAUNET_OK;
}
extern "C" int saunet_bilinear_bwd(const float* dy, int dy_ld, int B, int Hin, int Win, int C, float* dx, int dx_ld, int Hout, int Wout, int accumulate, void* stream) {
    SAUNET_CHECK_ARG(dy && dx && B > 0 && Hin > 0 && Win > 0 && C > 0 && Hout > 0 && Wout > 0, SAUNET_ERR_BAD_SHAPE, "bilinear_bwd: bad args");
    float sh = Hout > 1 ? (float)(Hin - 1) / (float)(Hout - 1) : 0.f, sw = Wout > 1 ? (float)(Win - 1) / (float)(Wout - 1) : 0.f;
    long long n = (long long)B * Hin * Win;
    if (C == 1 && (long long)Hout * Wout >= 16ll * Hin * Win) {
        bilinear_bwd_c1_warp_kernel<<<ew_blocks(n * 32), 256, 0, ST>>>(dy, dy_ld, B, Hin, Win, dx, dx_ld, Hout, Wout, sh, sw, accumulate);
        SAUNET_CHECK_LAUNCH("bilinear_bwd_c1_warp_kernel");
        return SAUNET_OK;
    }
    if (vec4(C, dy, dy_ld, dx, dx_ld)) bilinear_bwd_kernel<4><<<ew_blocks(n * (C / 4)), 256, 0, ST>>>(dy, dy_ld, B, Hin, Win, C, dx, dx_ld, Hout, Wout, sh, sw, accumulate);
    else bilinear_bwd_kernel<1><<<ew_blocks(n * C), 256, 0, ST>>>(dy, dy_ld, B, Hin, Win, C, dx, dx_ld, Hout, Wout, sh, sw, accumulate);
    SAUNET_CHECK_LAUNCH("bilinear_bwd_kernel");
    return SAUNET_OK;
}

template <bool MAXP>
static int pool_fwd(const float* x, int x_ld, int B, int Hin, int Win, int C, float* y, int y_ld, void* stream) {
    SAUNET_CHECK_ARG(x && y && B > 0 && Hin > 1 && Win > 1 && C > 0, SAUNET_ERR_BAD_SHAPE, "pool2_fwd: bad args");
    long long n = (long long)B * (Hin / 2) * (Win / 2);
    if (vec4(C, x, x_ld, y, y_ld)) pool2_fwd_kernel<4, MAXP><<<ew_blocks(n * (C / 4)), 256, 0, ST>>>(x, x_ld, B, Hin, Win, C, y, y_ld);
    else pool2_fwd_kernel<1, MAXP><<<ew_blocks(n * C), 256, 0, ST>>>(x, x_ld, B, Hin, Win, C, y, y_ld);
    SAUNET_CHECK_LAUNCH("pool2_fwd_kernel");
    return SAUNET_OK;
}
template <bool MAXP>
static int pool_bwd(const float* dy, int dy_ld, const float* x, int x_ld, int B, int Hin, int Win, int C, float* dx, int dx_ld, int accumulate, void* stream) {
    SAUNET_CHECK_ARG(dy && dx && (!MAXP || x) && B > 0 && Hin > 1 && Win > 1 && C > 0, SAUNET_ERR_BAD_SHAPE, "pool2_bwd: bad args");
    long long n = (long long)B * Hin * Win;
    if (vec4(C, dy, dy_ld, dx, dx_ld, x, x_ld)) pool2_bwd_kernel<4, MAXP><<<ew_blocks(n * (C / 4)), 256, 0, ST>>>(dy, dy_ld, x, x_ld, B, Hin, Win, C, dx, dx_ld, accumulate);
    else pool2_bwd_kernel<1, MAXP><<<ew_blocks(n * C), 256, 0, ST>>>(dy, dy_ld, x, x_ld, B, Hin, Win, C, dx, dx_ld, accumulate);
    SAUNET_CHECK_LAUNCH("pool2_bwd_kernel");
    return SAUNET_OK;
}
extern "C" int saunet_avgpool2_fwd(const float* x, int x_ld, int B, int Hin, int Win, int C, float* y, int y_ld, void* stream) { return pool_fwd<false>(x, x_ld, B, Hin, Win, C, y, y_ld, stream); }
extern "C" int saunet_maxpool2_fwd(const float* x, int x_ld, int B, int Hin, int Win, int C, float* y, int y_ld, void* stream) { return pool_fwd<true>(x, x_ld, B, Hin, Win, C, y, y_ld, stream); }
extern "C" int saunet_avgpool2_bwd(const float* dy, int dy_ld, int B, int Hin, int Win, int C, float* dx, int dx_ld, int accumulate, void* stream) { return pool_bwd<false>(dy, dy_ld, nullptr, 0, B, Hin, Win, C, dx, dx_ld, accumulate, stream); }
extern "C" int saunet_maxpool2_bwd(const float* dy, int dy_ld, const float* x, int x_ld, int B, int Hin, int Win, int C, float* dx, int dx_ld, int accumulate, void* stream) { return pool_bwd<true>(dy, dy_ld, x, x_ld, B, Hin, Win, C, dx, dx_ld, accumulate, stream); }

extern "C" int saunet_gap_fwd(const float* x, int x_ld, int B, long long HW, int C, float* y, void* stream) {
    SAUNET_CHECK_ARG(x && y && B > 0 && HW > 0 && C > 0, SAUNET_ERR_BAD_SHAPE, "gap_fwd: bad args");
    cudaError_t e = cudaMemsetAsync(y, 0, sizeof(float) * (size_t)B * C, ST);
    SAUNET_CHECK_ARG(e == cudaSuccess, SAUNET_ERR_CUDA, "gap_fwd: memset failed: %s", cudaGetErrorString(e));
    int cb = cdiv(C, 32);
    long long splits = (long long)kNumSMs * 4 / ((long long)cb * B); if (splits < 1) splits = 1;
    long long maxs = (HW + 63) / 64; if (splits > maxs) splits = maxs;
    if (C % 4 == 0 && x_ld % 4 == 0 && ((reinterpret_cast<uintptr_t>(x) & 15u) == 0)) {
        const int cb4 = (C / 4 + 31) / 32;
        long long sp = (long long)kNumSMs * 8 / ((long long)cb4 * B); if (sp < 1) sp = 1;
        long long mx = (HW + 255) / 256; if (sp > mx) sp = mx;
        gap_fwd4_kernel<<<dim3(cb4, B, (unsigned)sp), 256, 0, ST>>>(x, x_ld, HW, C, y, 1.0f / (float)HW);
        SAUNET_CHECK_LAUNCH("gap_fwd4_kernel");
        return SAUNET_OK;
    }
    gap_fwd_kernel<<<dim3(cb, B, (unsigned)splits), 256, 0, ST>>>(x, x_ld, HW, C, y, 1.0f / (float)HW);
    SAUNET_CHECK_LAUNCH("gap_fwd_kernel");
    return SAUNET_OK;
}
extern "C" int saunet_gap_bwd(const float* dy, int B, long long HW, int C, float* dx, int dx_ld, int accumulate, void* stream) {
    SAUNET_CHECK_ARG(dy && dx && B > 0 && HW > 0 && C > 0, SAUNET_ERR_BAD_SHAPE, "gap_bwd: bad args");
    long long n = (long long)B * HW;
    if (vec4(C, dy, C, dx, dx_ld)) gap_bwd_kernel<4><<<ew_blocks(n * (C / 4)), 256, 0, ST>>>(dy, B, HW, C, dx, dx_ld, accumulate, 1.0f / (float)HW);
    else gap_bwd_kernel<1><<<ew_blocks(n * C), 256, 0, ST>>>(dy, B, HW, C, dx, dx_ld, accumulate, 1.0f / (float)HW);
    SAUNET_CHECK_LAUNCH("gap_bwd_kernel");
    return SAUNET_OK;
}
extern "C" int saunet_act_bwd(const float* dy, int dy_ld, const float* y, int y_ld, float* dz, int dz_ld, int C, long long npix, int act, void* stream) {
    SAUNET_CHECK_ARG(dy && y && dz && C > 0 && npix > 0, SAUNET_ERR_BAD_SHAPE, "act_bwd: bad args");
    if (vec4(C, dy, dy_ld, y, y_ld, dz, dz_ld)) act_bwd_kernel<4><<<ew_blocks(npix * (C / 4)), 256, 0, ST>>>(dy, dy_ld, y, y_ld, dz, dz_ld, C, npix, act);
    else act_bwd_kernel<1><<<ew_blocks(npix * C), 256, 0, ST>>>(dy, dy_ld, y, y_ld, dz, dz_ld, C, npix, act);
    SAUNET_CHECK_LAUNCH("act_bwd_kernel");
    return SAUNET_OK;
}
extern "C" int saunet_dualatt_combine_fwd(const float* fused, int f_ld, const float* S, const float* cvec, int B, long long HW, int C, float* out, int o_ld, void* stream) {
    SAUNET_CHECK_ARG(fused && S && cvec && out && B > 0 && HW > 0 && C > 0, SAUNET_ERR_BAD_SHAPE, "dualatt_combine_fwd: bad args");
    long long npix = (long long)B * HW;
    if (vec4(C, fused, f_ld, out, o_ld, cvec, C)) dualatt_fwd_kernel<4><<<ew_blocks(npix * (C / 4)), 256, 0, ST>>>(fused, f_ld, S, cvec, HW, C, npix, out, o_ld);
    else dualatt_fwd_kernel<1><<<ew_blocks(npix * C), 256, 0, ST>>>(fused, f_ld, S, cvec, HW, C, npix, out, o_ld);
    SAUNET_CHECK_LAUNCH("dualatt_fwd_kernel");
    return SAUNET_OK;
}
extern "C" int saunet_dualatt_combine_bwd(const float* dout, int do_ld, const float* fused, int f_ld, const float* S, const float* cvec, int B, long long HW, int C,
                                          float* dfused, int df_ld, int df_acc, float* dS, float* dcvec, void* stream) {
    SAUNET_CHECK_ARG(dout && fused && S && cvec && dfused && dS && dcvec && B > 0 && HW > 0 && C > 0, SAUNET_ERR_BAD_SHAPE, "dualatt_combine_bwd: bad args");
    SAUNET_CHECK_ARG(C <= 1024, SAUNET_ERR_BAD_SHAPE, "dualatt_combine_bwd: C=%d > 1024 unsupported", C);
    long long chunks = (long long)kNumSMs * 4 / B; if (chunks < 1) chunks = 1;
    long long maxc = (HW + 31) / 32; if (chunks > maxc) chunks = maxc;
    dim3 grid((unsigned)chunks, B);
    if (C % 4 == 0 && C <= 512 && do_ld % 4 == 0 && f_ld % 4 == 0 && df_ld % 4 == 0 && aligned16(dout) && aligned16(fused) && aligned16(dfused) &&
        aligned16(cvec)) {
        if (C <= 128) dualatt_bwd4_kernel<1><<<grid, 256, 0, ST>>>(dout, do_ld, fused, f_ld, S, cvec, HW, C, dfused, df_ld, df_acc, dS, dcvec);
        else if (C <= 256) dualatt_bwd4_kernel<2><<<grid, 256, 0, ST>>>(dout, do_ld, fused, f_ld, S, cvec, HW, C, dfused, df_ld, df_acc, dS, dcvec);
        else dualatt_bwd4_kernel<4><<<grid, 256, 0, ST>>>(dout, do_ld, fused, f_ld, S, cvec, HW, C, dfused, df_ld, df_acc, dS, dcvec);
        SAUNET_CHECK_LAUNCH("dualatt_bwd4_kernel");
        return SAUNET_OK;
    }
    if (C <= 64) dualatt_bwd_kernel<2><<<grid, 256, 0, ST>>>(dout, do_ld, fused, f_ld, S, cvec, HW, C, dfused, df_ld, df_acc, dS, dcvec);
    else if (C <= 128) dualatt_bwd_kernel<4><<<grid, 256, 0, ST>>>(dout, do_ld, fused, f_ld, S, cvec, HW, C, dfused, df_ld, df_acc, dS, dcvec);
    else if (C <= 256) dualatt_bwd_kernel<8><<<grid, 256, 0, ST>>>(dout, do_ld, fused, f_ld, S, cvec, HW, C, dfused, df_ld, df_acc, dS, dcvec);
    else if (C <= 512) dualatt_bwd_kernel<16><<<grid, 256, 0, ST>>>(dout, do_ld, fused, f_ld, S, cvec, HW, C, dfused, df_ld, df_acc, dS, dcvec);
    else dualatt_bwd_kernel<32><<<grid, 256, 0, ST>>>(dout, do_ld, fused, f_ld, S, cvec, HW, C, dfused, df_ld, df_acc, dS, dcvec);
    SAUNET_CHECK_LAUNCH("dualatt_bwd_kernel");
    return SAUNET_OK;
}
extern "C" int saunet_rowscale_bwd(const float* dout, int do_ld, const float* out, int o_ld, const float* alpha, int C, long long npix, float* du, int du_ld, float* dalpha, int da_acc, void* stream) {
    SAUNET_CHECK_ARG(dout && out && alpha && du && dalpha && C > 0 && npix > 0, SAUNET_ERR_BAD_SHAPE, "rowscale_bwd: bad args");
    const int lpp = C / 4;
    if (C % 4 == 0 && (lpp == 1 || lpp == 2 || lpp == 4 || lpp == 8) && do_ld % 4 == 0 && o_ld % 4 == 0 && du_ld % 4 == 0 &&
        aligned16(dout) && aligned16(out) && aligned16(du)) {
        // grid-stride loop with the thread count a multiple of 32: every pixel's LPP lanes live in one warp and leave the loop together
        const int blocks = ew_blocks(npix * lpp);
        if (lpp == 8) rowscale_bwd4_kernel<8><<<blocks, 256, 0, ST>>>(dout, do_ld, out, o_ld, alpha, npix, du, du_ld, dalpha, da_acc);
        else if (lpp == 4) rowscale_bwd4_kernel<4><<<blocks, 256, 0, ST>>>(dout, do_ld, out, o_ld, alpha, npix, du, du_ld, dalpha, da_acc);
        else if (lpp == 2) rowscale_bwd4_kernel<2><<<blocks, 256, 0, ST>>>(dout, do_ld, out, o_ld, alpha, npix, du, du_ld, dalpha, da_acc);
        else rowscale_bwd4_kernel<1><<<blocks, 256, 0, ST>>>(dout, do_ld, out, o_ld, alpha, npix, du, du_ld, dalpha, da_acc);
        SAUNET_CHECK_LAUNCH("rowscale_bwd4_kernel");
        return SAUNET_OK;
    }
    rowscale_bwd_kernel<<<ew_blocks(npix * 32), 256, 0, ST>>>(dout, do_ld, out, o_ld, alpha, C, npix, du, du_ld, dalpha, da_acc);
    SAUNET_CHECK_LAUNCH("rowscale_bwd_kernel");
    return SAUNET_OK;
}
extern "C" int saunet_copy_slice(const float* src, int s_ld, float* dst, int d_ld, int C, long long npix, int accumulate, void* stream) {
    SAUNET_CHECK_ARG(src && dst && C > 0 && npix > 0 && s_ld >= C && d_ld >= C, SAUNET_ERR_BAD_SHAPE, "copy_slice: bad args");
    if (vec4(C, src, s_ld, dst, d_ld)) copy_slice_kernel<4><<<ew_blocks(npix * (C / 4)), 256, 0, ST>>>(src, s_ld, dst, d_ld, C, npix, accumulate);
    else copy_slice_kernel<1><<<ew_blocks(npix * C), 256, 0, ST>>>(src, s_ld, dst, d_ld, C, npix, accumulate);
    SAUNET_CHECK_LAUNCH("copy_slice_kernel");
    return SAUNET_OK;
}
extern "C" int saunet_nchw_to_nhwc(const float* src, float* dst, int dst_ld, int B, int C, long long HW, void* stream) {
    SAUNET_CHECK_ARG(src && dst && B > 0 && C > 0 && HW > 0 && dst_ld >= C && B <= 65535, SAUNET_ERR_BAD_SHAPE, "nchw_to_nhwc: bad args");
    transpose_kernel<true><<<dim3(cdiv(HW, 32), cdiv(C, 32), B), 256, 0, ST>>>(src, dst, dst_ld, C, HW);
    SAUNET_CHECK_LAUNCH("transpose_kernel");
    return SAUNET_OK;
}
extern "C" int saunet_nhwc_to_nchw(const float* src, int src_ld, float* dst, int B, int C, long long HW, void* stream) {
    SAUNET_CHECK_ARG(src && dst && B > 0 && C > 0 && HW > 0 && src_ld >= C && B <= 65535, SAUNET_ERR_BAD_SHAPE, "nhwc_to_nchw: bad args");
    transpose_kernel<false><<<dim3(cdiv(HW, 32), cdiv(C, 32), B), 256, 0, ST>>>(src, dst, src_ld, C, HW);
    SAUNET_CHECK_LAUNCH("transpose_kernel");
    return SAUNET_OK;
}
