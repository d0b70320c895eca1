// conv_simt.cu -- fp32 FFMA implicit-GEMM convolution (forward / data-grad via gather-GEMM,
// weight-grad via pixel-split outer-product GEMM) plus weight (un)packing.
//
// This is the exact-fp32 path: it serves every conv geometry on the SAUNet path (odd channel
// counts such as the 3-channel stem, the C+1 gate convs and the 1-channel heads) and is the
// cross-check for the tcgen05 path in conv_tc.cu.  NHWC activations, K ordered (ky,kx,c).
#include "common.cuh"

namespace saunet {

constexpr int BM = 128;   // output pixels per CTA
constexpr int BK = 16;    // K slice
constexpr int APAD = 4;

struct ConvP {
    saunet_conv_desc d;
    int M, K, HgWg;
};

template <int BN, int TN, bool VA, bool VB>
__global__ void __launch_bounds__(16 * (BN / TN)) conv_fwd_kernel(const ConvP p) {
    constexpr int NTX = BN / TN;
    constexpr int NT = 16 * NTX;
    constexpr int BPAD = 4;
    __shared__ __align__(16) float As[2][BK][BM + APAD];
    __shared__ __align__(16) float Bs[2][BK][BN + BPAD];
    const saunet_conv_desc& d = p.d;
    const int tid = threadIdx.x;
    const int tx = tid % NTX, ty = tid / NTX;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

    // ---- A (im2col gather) load bookkeeping
    constexpr int NA = VA ? (BM * BK / 4) / NT : (BM * BK) / NT;
    constexpr int NB = VB ? (BK * BN / 4 + NT - 1) / NT : (BK * BN + NT - 1) / NT;
    int a_iy0[VA ? NA : 1], a_ix0[VA ? NA : 1], a_b[VA ? NA : 1];
    if constexpr (VA) {
#pragma unroll
        for (int s = 0; s < NA; ++s) {
            int row = (tid >> 2) + s * (NT / 4);
            int m = m0 + row;
            if (m < p.M) {
                int b = m / p.HgWg; int r = m - b * p.HgWg; int i = r / d.Wg; int j = r - i * d.Wg;
                a_b[s] = b; a_iy0[s] = i * d.sy + d.offy; a_ix0[s] = j * d.sx + d.offx;
            } else { a_b[s] = -1; a_iy0[s] = 0; a_ix0[s] = 0; }
        }
    }
    float4 ra4[VA ? NA : 1];
    float ra1[VA ? 1 : NA];
    float4 rb4[VB ? NB : 1];
    float rb1[VB ? 1 : NB];

    auto load_a = [&](int kt) {
        if constexpr (VA) {
            const int k = kt * BK + (tid & 3) * 4;
            int c = 0, ky = 0, kx = 0; bool kval = k < p.K;
            if (kval) { int tap = k / d.Cin; c = k - tap * d.Cin; ky = tap / d.KW; kx = tap - ky * d.KW; }
            float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
            if (d.in_scale && kval) {
                sc = *reinterpret_cast<const float4*>(d.in_scale + c);
                sh = *reinterpret_cast<const float4*>(d.in_shift + c);
            }
#pragma unroll
            for (int s = 0; s < NA; ++s) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                int iy = a_iy0[s] + ky, ix = a_ix0[s] + kx;
                if (kval && a_b[s] >= 0 && iy >= 0 && iy < d.Hin && ix >= 0 && ix < d.Win) {
                    const float* ptr = d.x + ((size_t)(a_b[s] * d.Hin + iy) * d.Win + ix) * d.x_ld + c;
                    v = __ldg(reinterpret_cast<const float4*>(ptr));
                    if (d.in_scale) {
                        v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y);
                        v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
                        if (d.in_relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                    }
                }
                ra4[s] = v;
            }
        } else {
            const int k = kt * BK + (tid & 15);
            int c = 0, ky = 0, kx = 0; bool kval = k < p.K;
            if (kval) { int tap = k / d.Cin; c = k - tap * d.Cin; ky = tap / d.KW; kx = tap - ky * d.KW; }
            float sc = 1.f, sh = 0.f;
            if (d.in_scale && kval) { sc = d.in_scale[c]; sh = d.in_shift[c]; }
#pragma unroll
            for (int s = 0; s < NA; ++s) {
                int row = (tid >> 4) + s * (NT / 16);
                int m = m0 + row;
                float v = 0.f;
                if (kval && m < p.M) {
                    int b = m / p.HgWg; int r = m - b * p.HgWg; int i = r / d.Wg; int j = r - i * d.Wg;
                    int iy = i * d.sy + d.offy + ky, ix = j * d.sx + d.offx + kx;
                    if (iy >= 0 && iy < d.Hin && ix >= 0 && ix < d.Win) {
                        v = __ldg(d.x + ((size_t)(b * d.Hin + iy) * d.Win + ix) * d.x_ld + c);
                        if (d.in_scale) { v = fmaf(v, sc, sh); if (d.in_relu) v = fmaxf(v, 0.f); }
                    }
                }
                ra1[s] = v;
            }
        }
    };
    auto load_b = [&](int kt) {
        if constexpr (VB) {
#pragma unroll
            for (int s = 0; s < NB; ++s) {
                int idx = tid + s * NT;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (idx < BK * BN / 4) {
                    int kr = idx / (BN / 4), nq = idx - kr * (BN / 4);
                    int k = kt * BK + kr, n = n0 + nq * 4;
                    if (k < p.K && n < d.Cout) v = __ldg(reinterpret_cast<const float4*>(d.w + (size_t)k * d.Cout + n));
                }
                rb4[s] = v;
            }
        } else {
#pragma unroll
            for (int s = 0; s < NB; ++s) {
                int idx = tid + s * NT;
                float v = 0.f;
                if (idx < BK * BN) {
                    int kr = idx / BN, nn = idx - kr * BN;
                    int k = kt * BK + kr, n = n0 + nn;
                    if (k < p.K && n < d.Cout) v = __ldg(d.w + (size_t)k * d.Cout + n);
                }
                rb1[s] = v;
            }
        }
    };
    auto store_ab = [&](int buf) {
        if constexpr (VA) {
            const int kq = (tid & 3) * 4;
#pragma unroll
            for (int s = 0; s < NA; ++s) {
                int row = (tid >> 2) + s * (NT / 4);
                As[buf][kq + 0][row] = ra4[s].x; As[buf][kq + 1][row] = ra4[s].y;
                As[buf][kq + 2][row] = ra4[s].z; As[buf][kq + 3][row] = ra4[s].w;
            }
        } else {
#pragma unroll
            for (int s = 0; s < NA; ++s) As[buf][tid & 15][(tid >> 4) + s * (NT / 16)] = ra1[s];
        }
        if constexpr (VB) {
#pragma unroll
            for (int s = 0; s < NB; ++s) {
                int idx = tid + s * NT;
                if (idx < BK * BN / 4) {
                    int kr = idx / (BN / 4), nq = idx - kr * (BN / 4);
                    *reinterpret_cast<float4*>(&Bs[buf][kr][nq * 4]) = rb4[s];
                }
            }
        } else {
#pragma unroll
            for (int s = 0; s < NB; ++s) {
                int idx = tid + s * NT;
                if (idx < BK * BN) { int kr = idx / BN, nn = idx - kr * BN; Bs[buf][kr][nn] = rb1[s]; }
            }
        }
    };

    float acc[8][TN];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int e = 0; e < TN; ++e) acc[r][e] = 0.f;

    const int nkt = (p.K + BK - 1) / BK;
    load_a(0); load_b(0); store_ab(0);
    __syncthreads();
    for (int kt = 0; kt < nkt; ++kt) {
        const int cur = kt & 1;
        if (kt + 1 < nkt) { load_a(kt + 1); load_b(kt + 1); }
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[8], b[TN];
            float4 a0 = *reinterpret_cast<const float4*>(&As[cur][k][ty * 4]);
            float4 a1 = *reinterpret_cast<const float4*>(&As[cur][k][64 + ty * 4]);
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
            if constexpr (TN == 8) {
                float4 b0 = *reinterpret_cast<const float4*>(&Bs[cur][k][tx * 4]);
                float4 b1 = *reinterpret_cast<const float4*>(&Bs[cur][k][BN / 2 + tx * 4]);
                b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
            } else if constexpr (TN == 4) {
                float4 b0 = *reinterpret_cast<const float4*>(&Bs[cur][k][tx * 4]);
                b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
            } else if constexpr (TN == 2) {
                float2 b0 = *reinterpret_cast<const float2*>(&Bs[cur][k][tx * 2]);
                b[0] = b0.x; b[1] = b0.y;
            } else {
                b[0] = Bs[cur][k][tx];
            }
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int e = 0; e < TN; ++e) acc[r][e] = fmaf(a[r], b[e], acc[r][e]);
        }
        if (kt + 1 < nkt) store_ab(cur ^ 1);
        __syncthreads();
    }

    // ---- epilogue
    int col[TN];
#pragma unroll
    for (int e = 0; e < TN; ++e) col[e] = n0 + ((TN == 8) ? (e < 4 ? tx * 4 + e : BN / 2 + tx * 4 + (e - 4)) : tx * TN + e);
    float bias[TN];
#pragma unroll
    for (int e = 0; e < TN; ++e) bias[e] = (d.bias && col[e] < d.Cout) ? d.bias[col[e]] : 0.f;
    float ssum[TN], ssq[TN];
#pragma unroll
    for (int e = 0; e < TN; ++e) { ssum[e] = 0.f; ssq[e] = 0.f; }
    const bool vst = (d.Cout % 4 == 0) && (d.y_ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(d.y) & 15u) == 0);
    const bool ident = (d.osy == 1 && d.osx == 1 && d.oy0 == 0 && d.ox0 == 0 && d.Hout == d.Hg && d.Wout == d.Wg);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int m = m0 + (r < 4 ? ty * 4 + r : 64 + ty * 4 + (r - 4));
        if (m >= p.M) continue;
        size_t opix;
        if (ident) opix = (size_t)m;
        else {
            int b = m / p.HgWg; int rr = m - b * p.HgWg; int i = rr / d.Wg; int j = rr - i * d.Wg;
            opix = (size_t)(b * d.Hout + i * d.osy + d.oy0) * d.Wout + (j * d.osx + d.ox0);
        }
        float* yp = d.y + opix * d.y_ld;
        const float rs = d.row_scale ? (d.row_scale[m] + d.row_scale_add) : 1.f;
        float o[TN];
#pragma unroll
        for (int e = 0; e < TN; ++e) {
            float v = acc[r][e] + bias[e];
            if (col[e] < d.Cout) { ssum[e] += v; ssq[e] = fmaf(v, v, ssq[e]); }
            o[e] = apply_act(v * rs, d.act);
        }
        if (TN >= 4 && vst) {
#pragma unroll
            for (int h = 0; h < TN / 4; ++h) {
                if (col[h * 4] < d.Cout) {
                    float4* dst = reinterpret_cast<float4*>(yp + col[h * 4]);
                    float4 v = make_float4(o[h * 4 + 0], o[h * 4 + 1], o[h * 4 + 2], o[h * 4 + 3]);
                    if (d.accumulate) { float4 c = *dst; v.x += c.x; v.y += c.y; v.z += c.z; v.w += c.w; }
                    *dst = v;
                }
            }
        } else {
#pragma unroll
            for (int e = 0; e < TN; ++e) {
                if (col[e] < d.Cout) {
                    float v = o[e];
                    if (d.accumulate) v += yp[col[e]];
                    yp[col[e]] = v;
                }
            }
        }
    }
    if (d.stat_sum) {
        float* red = &As[0][0][0];      // reuse: [2][BN]
        __syncthreads();
        for (int i = tid; i < 2 * BN; i += NT) red[i] = 0.f;
        __syncthreads();
#pragma unroll
        for (int e = 0; e < TN; ++e) {
            int lc = col[e] - n0;
            atomicAdd(&red[lc], ssum[e]);
            atomicAdd(&red[BN + lc], ssq[e]);
        }
        __syncthreads();
        for (int i = tid; i < BN; i += NT) {
            if (n0 + i < d.Cout) {
                atomicAdd(d.stat_sum + n0 + i, (double)red[i]);
                atomicAdd(d.stat_sumsq + n0 + i, (double)red[BN + i]);
            }
        }
    }
}

template <int BN, int TN>
static int launch_fwd(const ConvP& p, bool va, bool vb, cudaStream_t st) {
    dim3 grid(cdiv(p.M, BM), cdiv(p.d.Cout, BN));
    dim3 block(16 * (BN / TN));
    if (va && vb) conv_fwd_kernel<BN, TN, true, true><<<grid, block, 0, st>>>(p);
    else if (va) conv_fwd_kernel<BN, TN, true, false><<<grid, block, 0, st>>>(p);
    else if (vb) conv_fwd_kernel<BN, TN, false, true><<<grid, block, 0, st>>>(p);
    else conv_fwd_kernel<BN, TN, false, false><<<grid, block, 0, st>>>(p);
    SAUNET_CHECK_LAUNCH("conv_fwd_kernel");
    return SAUNET_OK;
}

int conv_fwd_simt(const saunet_conv_desc* d, cudaStream_t st) {
    ConvP p; p.d = *d;
    long long M = (long long)d->B * d->Hg * d->Wg;
    SAUNET_CHECK_ARG(M > 0 && M < (1ll << 31), SAUNET_ERR_BAD_SHAPE, "conv2d_fwd: bad M=%lld", M);
    p.M = (int)M; p.K = d->KH * d->KW * d->Cin; p.HgWg = d->Hg * d->Wg;
    bool va = (d->Cin % 4 == 0) && (d->x_ld % 4 == 0) && aligned16(d->x) &&
              (!d->in_scale || (aligned16(d->in_scale) && aligned16(d->in_shift)));
    bool vb = (d->Cout % 4 == 0) && aligned16(d->w);
    if (d->Cout >= 128) return launch_fwd<128, 8>(p, va, vb, st);
    if (d->Cout > 32) return launch_fwd<64, 8>(p, va, vb, st);
    if (d->Cout > 16) return launch_fwd<32, 4>(p, va, vb, st);
    return launch_fwd<16, 2>(p, va, vb, st);
}

// ------------------------------------------------------------------------------------------------
// weight gradient
struct WgP {
    saunet_wgrad_desc d;
    long long M;
    int HgWg;
    long long pix_per_split;
};

constexpr int WBA = 64, WBB = 64, WBK = 16;

template <bool VP, bool VQ>
__global__ void __launch_bounds__(128) conv_wgrad_kernel(const WgP p) {
    __shared__ __align__(16) float Ps[2][WBK][WBA + 4];
    __shared__ __align__(16) float Qs[2][WBK][WBB + 4];
    const saunet_wgrad_desc& d = p.d;
    const int tid = threadIdx.x;
    const int ta = tid & 7, tb = tid >> 3;          // a rows: ta*4.., 32+ta*4.. ; b cols: tb*4..
    const int a0 = blockIdx.x * WBA;
    const int nbt = (d.Cb + WBB - 1) / WBB;
    const int tap = blockIdx.y / nbt;
    const int b0 = (blockIdx.y - tap * nbt) * WBB;
    const int ky = tap / d.KW, kx = tap - ky * d.KW;
    const long long mbeg = (long long)blockIdx.z * p.pix_per_split;
    long long mend = mbeg + p.pix_per_split; if (mend > p.M) mend = p.M;

    constexpr int NP = VP ? 2 : 8;
    constexpr int NQ = VQ ? 2 : 8;
    float4 rp4[VP ? NP : 1]; float rp1[VP ? 1 : NP];
    float4 rq4[VQ ? NQ : 1]; float rq1[VQ ? 1 : NQ];

    auto load = [&](long long mb) {
        if constexpr (VP) {
            const int a = a0 + (tid & 15) * 4;
#pragma unroll
            for (int s = 0; s < NP; ++s) {
                long long m = mb + (tid >> 4) + s * 8;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (m < mend && a < d.Ca) v = __ldg(reinterpret_cast<const float4*>(d.p + (size_t)m * d.p_ld + a));
                rp4[s] = v;
            }
        } else {
            const int a = a0 + (tid & 63);
#pragma unroll
            for (int s = 0; s < NP; ++s) {
                long long m = mb + (tid >> 6) + s * 2;
                rp1[s] = (m < mend && a < d.Ca) ? __ldg(d.p + (size_t)m * d.p_ld + a) : 0.f;
            }
        }
        if constexpr (VQ) {
            const int b = b0 + (tid & 15) * 4;
            float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
            if (d.q_scale && b < d.Cb) {
                sc = *reinterpret_cast<const float4*>(d.q_scale + b);
                sh = *reinterpret_cast<const float4*>(d.q_shift + b);
            }
#pragma unroll
            for (int s = 0; s < NQ; ++s) {
                long long m = mb + (tid >> 4) + s * 8;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (m < mend && b < d.Cb) {
                    int bi = (int)(m / p.HgWg); int r = (int)(m - (long long)bi * p.HgWg); int i = r / d.Wg; int j = r - i * d.Wg;
                    int iy = i * d.sy + ky + d.offy, ix = j * d.sx + kx + d.offx;
                    if (iy >= 0 && iy < d.Hq && ix >= 0 && ix < d.Wq) {
                        v = __ldg(reinterpret_cast<const float4*>(d.q + ((size_t)(bi * d.Hq + iy) * d.Wq + ix) * d.q_ld + b));
                        if (d.q_scale) {
                            v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y);
                            v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
                            if (d.q_relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                        }
                    }
                }
                rq4[s] = v;
            }
        } else {
            const int b = b0 + (tid & 63);
            float sc = 1.f, sh = 0.f;
            if (d.q_scale && b < d.Cb) { sc = d.q_scale[b]; sh = d.q_shift[b]; }
#pragma unroll
            for (int s = 0; s < NQ; ++s) {
                long long m = mb + (tid >> 6) + s * 2;
                float v = 0.f;
                if (m < mend && b < d.Cb) {
                    int bi = (int)(m / p.HgWg); int r = (int)(m - (long long)bi * p.HgWg); int i = r / d.Wg; int j = r - i * d.Wg;
                    int iy = i * d.sy + ky + d.offy, ix = j * d.sx + kx + d.offx;
                    if (iy >= 0 && iy < d.Hq && ix >= 0 && ix < d.Wq) {
                        v = __ldg(d.q + ((size_t)(bi * d.Hq + iy) * d.Wq + ix) * d.q_ld + b);
                        if (d.q_scale) { v = fmaf(v, sc, sh); if (d.q_relu) v = fmaxf(v, 0.f); }
                    }
                }
                rq1[s] = v;
            }
        }
    };
    auto store = [&](int buf) {
        if constexpr (VP) {
#pragma unroll
            for (int s = 0; s < NP; ++s) *reinterpret_cast<float4*>(&Ps[buf][(tid >> 4) + s * 8][(tid & 15) * 4]) = rp4[s];
        } else {
#pragma unroll
            for (int s = 0; s < NP; ++s) Ps[buf][(tid >> 6) + s * 2][tid & 63] = rp1[s];
        }
        if constexpr (VQ) {
#pragma unroll
            for (int s = 0; s < NQ; ++s) *reinterpret_cast<float4*>(&Qs[buf][(tid >> 4) + s * 8][(tid & 15) * 4]) = rq4[s];
        } else {
#pragma unroll
            for (int s = 0; s < NQ; ++s) Qs[buf][(tid >> 6) + s * 2][tid & 63] = rq1[s];
        }
    };

    float acc[8][4];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[r][e] = 0.f;

    const int nchunks = (int)((mend - mbeg + WBK - 1) / WBK);
    if (nchunks > 0) {
        load(mbeg); store(0);
        __syncthreads();
        for (int ch = 0; ch < nchunks; ++ch) {
            const int cur = ch & 1;
            if (ch + 1 < nchunks) load(mbeg + (long long)(ch + 1) * WBK);
#pragma unroll
            for (int k = 0; k < WBK; ++k) {
                float4 x0 = *reinterpret_cast<const float4*>(&Ps[cur][k][ta * 4]);
                float4 x1 = *reinterpret_cast<const float4*>(&Ps[cur][k][32 + ta * 4]);
                float4 yb = *reinterpret_cast<const float4*>(&Qs[cur][k][tb * 4]);
                float a[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
                float b[4] = {yb.x, yb.y, yb.z, yb.w};
#pragma unroll
                for (int r = 0; r < 8; ++r)
#pragma unroll
                    for (int e = 0; e < 4; ++e) acc[r][e] = fmaf(a[r], b[e], acc[r][e]);
            }
            if (ch + 1 < nchunks) store(cur ^ 1);
            __syncthreads();
        }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int b = b0 + tb * 4 + e;
        if (b >= d.Cb) continue;
        float* row = d.dw + ((size_t)tap * d.Cb + b) * d.Ca;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int a = a0 + (r < 4 ? ta * 4 + r : 32 + ta * 4 + (r - 4));
            if (a < d.Ca) atomicAdd(row + a, acc[r][e]);
        }
    }
}

int conv_wgrad_simt(const saunet_wgrad_desc* d, cudaStream_t st) {
    WgP p; p.d = *d;
    p.M = (long long)d->B * d->Hg * d->Wg; p.HgWg = d->Hg * d->Wg;
    SAUNET_CHECK_ARG(p.M > 0, SAUNET_ERR_BAD_SHAPE, "conv2d_wgrad: empty problem");
    const int taps = d->KH * d->KW;
    const int gx = cdiv(d->Ca, WBA), gy = taps * cdiv(d->Cb, WBB);
    // pixel splits: aim for >= 4 waves of 148 SMs x 4 CTAs, at least 256 pixels per split
    long long want = (long long)kNumSMs * 8 / ((long long)gx * gy) + 1;
    long long maxs = (p.M + 255) / 256;
    long long splits = want < 1 ? 1 : (want > maxs ? maxs : want);
    if (splits > 65535) splits = 65535;
    long long pps = (p.M + splits - 1) / splits;
    pps = (pps + WBK - 1) / WBK * WBK;
    splits = (p.M + pps - 1) / pps;
    p.pix_per_split = pps;
    bool vp = (d->Ca % 4 == 0) && (d->p_ld % 4 == 0) && aligned16(d->p);
    bool vq = (d->Cb % 4 == 0) && (d->q_ld % 4 == 0) && aligned16(d->q) &&
              (!d->q_scale || (aligned16(d->q_scale) && aligned16(d->q_shift)));
    dim3 grid(gx, gy, (unsigned)splits);
    if (vp && vq) conv_wgrad_kernel<true, true><<<grid, 128, 0, st>>>(p);
    else if (vp) conv_wgrad_kernel<true, false><<<grid, 128, 0, st>>>(p);
    else if (vq) conv_wgrad_kernel<false, true><<<grid, 128, 0, st>>>(p);
    else conv_wgrad_kernel<false, false><<<grid, 128, 0, st>>>(p);
    SAUNET_CHECK_LAUNCH("conv_wgrad_kernel");
    return SAUNET_OK;
}

// ------------------------------------------------------------------------------------------------
// weight packing.  w is [A][Bc][T] (T = KH*KW taps, row-major ky,kx)
__global__ void pack_weights_kernel(const float* __restrict__ w, float* __restrict__ packed, int A, int Bc, int KH, int KW, int mode) {
    const int T = KH * KW;
    const long long n = (long long)A * Bc * T;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
        if (mode == 0) {            // packed[(t,b)][a] = w[a][b][t]
            int a = (int)(idx % A); long long r = idx / A; int b = (int)(r % Bc); int t = (int)(r / Bc);
            packed[idx] = w[((size_t)a * Bc + b) * T + t];
        } else if (mode == 1) {     // packed[(t,a)][b] = w[a][b][T-1-t]
            int b = (int)(idx % Bc); long long r = idx / Bc; int a = (int)(r % A); int t = (int)(r / A);
            packed[idx] = w[((size_t)a * Bc + b) * T + (T - 1 - t)];
        } else {                    // packed[ph][(ty,tx,a)][b] = w[a][b][3-pa-2ty][3-pb-2tx]
            int b = (int)(idx % Bc); long long r = idx / Bc; int a = (int)(r % A); r /= A;
            int tx = (int)(r % 2); r /= 2; int ty = (int)(r % 2); int ph = (int)(r / 2);
            int pa = ph >> 1, pb = ph & 1;
            int ky = 3 - pa - 2 * ty, kx = 3 - pb - 2 * tx;
            packed[idx] = w[((size_t)a * Bc + b) * 16 + ky * 4 + kx];
        }
    }
}
__global__ void unpack_wgrad_kernel(const float* __restrict__ packed, float* __restrict__ wg, int A, int Bc, int T, int accumulate) {
    const long long n = (long long)A * Bc * T;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
        int t = (int)(idx % T); long long r = idx / T; int b = (int)(r % Bc); int a = (int)(r / Bc);
        float v = packed[((size_t)t * Bc + b) * A + a];
        wg[idx] = accumulate ? wg[idx] + v : v;
    }
}

// all conv weight gradients of a model in ONE launch: entry j describes parameter j (packed [(t,b)][a] image at
// packed_base + packed_off, parameter-layout gradient w[a][b][t] at grad_base + grad_off); `first` = prefix sum of
// element counts.  Thread -> flat element index -> binary search for its parameter.
__global__ void unpack_wgrad_multi_kernel(const saunet_unpack_entry* __restrict__ tab, int n, long long total,
                                          float* __restrict__ packed_base, float* __restrict__ grad_base) {
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        int lo = 0, hi = n - 1;
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (tab[mid].first <= idx) lo = mid; else hi = mid - 1; }
        const saunet_unpack_entry e = tab[lo];
        const long long loc = idx - e.first;
        const int t = (int)(loc % e.T); const long long r = loc / e.T; const int b = (int)(r % e.Bc); const int a = (int)(r / e.Bc);
        float* src = packed_base + e.packed_off + ((size_t)t * e.Bc + b) * e.A + a;
        grad_base[e.grad_off + loc] += *src;
        *src = 0.f;                                    // consumed: the image is all-zero again for the next backward
    }
}

}  // namespace saunet

using namespace saunet;

extern "C" int saunet_unpack_wgrad_multi(const saunet_unpack_entry* table, int n, long long total, float* packed_base,
                                         float* grad_base, void* stream) {
    SAUNET_CHECK_ARG(table && packed_base && grad_base && n > 0 && total > 0, SAUNET_ERR_BAD_SHAPE, "unpack_wgrad_multi: bad args");
    long long blocks = (total + 255) / 256; if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    unpack_wgrad_multi_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(table, n, total, packed_base, grad_base);
    SAUNET_CHECK_LAUNCH("unpack_wgrad_multi_kernel");
    return SAUNET_OK;
}

extern "C" int saunet_pack_weights(const float* w, float* packed, int A, int Bc, int KH, int KW, int mode, void* stream) {
    SAUNET_CHECK_ARG(w && packed && A > 0 && Bc > 0 && KH > 0 && KW > 0, SAUNET_ERR_BAD_SHAPE, "pack_weights: bad args");
    SAUNET_CHECK_ARG(mode >= 0 && mode <= 2 && (mode != 2 || (KH == 4 && KW == 4)), SAUNET_ERR_BAD_SHAPE, "pack_weights: bad mode");
    long long n = (long long)A * Bc * KH * KW;
    int blocks = (int)((n + 255) / 256); if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    pack_weights_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w, packed, A, Bc, KH, KW, mode);
    SAUNET_CHECK_LAUNCH("pack_weights_kernel");
    return SAUNET_OK;
}
extern "C" int saunet_unpack_wgrad(const float* packed, float* wgrad, int A, int Bc, int KH, int KW, int accumulate, void* stream) {
    SAUNET_CHECK_ARG(wgrad && packed && A > 0 && Bc > 0 && KH > 0 && KW > 0, SAUNET_ERR_BAD_SHAPE, "unpack_wgrad: bad args");
    long long n = (long long)A * Bc * KH * KW;
    int blocks = (int)((n + 255) / 256); if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    unpack_wgrad_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(packed, wgrad, A, Bc, KH * KW, accumulate);
    SAUNET_CHECK_LAUNCH("unpack_wgrad_kernel");
    return SAUNET_OK;
}
