// conv_skinny.cu -- 1x1 convolutions with very few channels on very many pixels (the gated shape stream at full
// resolution: C+1 -> C+1 and C+1 -> 1 gate convs, fuse 8->1, cw 2->1, expand 1->32, final 32->4 and their data /
// weight gradients; models/GSConv.py:38-45, models/models.py:291-301,324).  These are HBM-bound (a few hundred bytes
// per pixel, <= 40x40 MACs): one thread per pixel, weights broadcast from shared memory, no tensor cores, any
// channel count / alignment.  Same descriptor and the same fused prologue / epilogue as the GEMM kernels.
#include "common.cuh"
#include <stdlib.h>

namespace saunet {

constexpr int kSkMaxC = 40;

struct SkP {
    saunet_conv_desc d;
    int M;
};

// ---- forward / data gradient: y[m][0:Cout) = act(rs[m] * (sum_c pro(x[m][c]) * W[c][n] + bias[n])) ----
// 128 pixels per block iteration: the input rows are copied to shared memory with coalesced loads (one thread per
// pixel reading its own 100+ byte row straight from global would touch 32 cache lines per instruction), each
// thread then computes its pixel from a conflict-free (odd-stride) shared row, and the outputs go back through
// the same buffer so the global stores are coalesced too.
constexpr int kSkTile = 128;

template <int CO>
__global__ void __launch_bounds__(kSkTile) skinny_fwd_kernel(const SkP p) {
    __shared__ __align__(16) float Ws[kSkMaxC * CO];
    __shared__ float bs[CO], scs[kSkMaxC], shs[kSkMaxC];
    __shared__ float tile[kSkTile * (kSkMaxC + 1)];
    __shared__ float red[kSkTile / 32][2 * CO];
    const saunet_conv_desc& d = p.d;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int Cin = d.Cin, Cout = d.Cout;
    const int si = Cin | 1, so = Cout | 1;                    // odd row strides: no bank conflicts
    for (int i = tid; i < Cin * CO; i += kSkTile) {
        const int c = i / CO, n = i - c * CO;
        Ws[i] = n < Cout ? d.w[(size_t)c * Cout + n] : 0.f;
    }
    for (int i = tid; i < CO; i += kSkTile) bs[i] = (d.bias && i < Cout) ? d.bias[i] : 0.f;
    for (int i = tid; i < Cin; i += kSkTile) { scs[i] = d.in_scale ? d.in_scale[i] : 1.f; shs[i] = d.in_scale ? d.in_shift[i] : 0.f; }
    float ssum[CO], ssq[CO];
#pragma unroll
    for (int n = 0; n < CO; ++n) { ssum[n] = 0.f; ssq[n] = 0.f; }
    for (int m0 = blockIdx.x * kSkTile; m0 < p.M; m0 += gridDim.x * kSkTile) {
        const int npx = min(kSkTile, p.M - m0);
        __syncthreads();
        for (int i = tid; i < npx * Cin; i += kSkTile) {
            const int px = i / Cin, c = i - px * Cin;
            tile[px * si + c] = __ldg(d.x + (size_t)(m0 + px) * d.x_ld + c);
        }
        __syncthreads();
        float acc[CO];
#pragma unroll
        for (int n = 0; n < CO; ++n) acc[n] = bs[n];
        const bool valid = tid < npx;
        if (valid) {
            const float* xr = tile + tid * si;
            for (int c = 0; c < Cin; ++c) {
                float xv = xr[c];
                if (d.in_scale) { xv = fmaf(xv, scs[c], shs[c]); if (d.in_relu) xv = fmaxf(xv, 0.f); }
                const float* wr = Ws + c * CO;
#pragma unroll
                for (int n = 0; n < CO; ++n) acc[n] = fmaf(xv, wr[n], acc[n]);
            }
        }
        __syncthreads();                                      // everyone is done reading the input tile
        if (valid) {
            const float rs = d.row_scale ? (d.row_scale[m0 + tid] + d.row_scale_add) : 1.f;
#pragma unroll
            for (int n = 0; n < CO; ++n) {
                if (n < Cout) {
                    ssum[n] += acc[n]; ssq[n] = fmaf(acc[n], acc[n], ssq[n]);
                    tile[tid * so + n] = apply_act(acc[n] * rs, d.act);
                }
            }
        }
        __syncthreads();
        for (int i = tid; i < npx * Cout; i += kSkTile) {
            const int px = i / Cout, n = i - px * Cout;
            float* yp = d.y + (size_t)(m0 + px) * d.y_ld + n;
            const float o = tile[px * so + n];
            *yp = d.accumulate ? *yp + o : o;
        }
    }
    if (d.stat_sum) {
#pragma unroll
        for (int n = 0; n < CO; ++n) {
            const float a = warp_sum(ssum[n]), b = warp_sum(ssq[n]);
            if (lane == 0) { red[warp][n] = a; red[warp][CO + n] = b; }
        }
        __syncthreads();
        for (int i = tid; i < 2 * CO; i += kSkTile) {
            const int n = i % CO;
            if (n < Cout) {
                float s = 0.f;
#pragma unroll
                for (int w = 0; w < kSkTile / 32; ++w) s += red[w][i];
                atomicAdd((i < CO ? d.stat_sum : d.stat_sumsq) + n, (double)s);
            }
        }
    }
}


// ---- forward, row-per-thread variant (16-byte aligned rows) -------------------------------------------------------
// One thread = one pixel: its input row is read with up to 10 independent 128-bit loads straight into registers (all
// in flight at once), the [Cin][CO] weights are broadcast from shared memory four at a time (LDS.128 : FFMA = 1 : 4)
// and the output row is written with 128-bit stores.  No shared-memory staging of activations, no block barriers in
// the pixel loop.  (The staged kernel above is kept for unaligned rows.)
template <int CO, int CI4, bool STATS>
__global__ void __launch_bounds__(kSkTile) skinny_fwd_row_kernel(const SkP p) {
    __shared__ __align__(16) float Ws[4 * CI4 * CO];
    __shared__ __align__(16) float bs[CO];
    __shared__ float scs[4 * CI4], shs[4 * CI4];
    __shared__ float red[kSkTile / 32][2 * CO];
    const saunet_conv_desc& d = p.d;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int Cin = d.Cin, Cout = d.Cout;
    for (int i = tid; i < 4 * CI4 * CO; i += kSkTile) {
        const int c = i / CO, n = i - c * CO;
        Ws[i] = (c < Cin && n < Cout) ? d.w[(size_t)c * Cout + n] : 0.f;
    }
    for (int i = tid; i < CO; i += kSkTile) bs[i] = (d.bias && i < Cout) ? d.bias[i] : 0.f;
    for (int i = tid; i < 4 * CI4; i += kSkTile) {
        scs[i] = (d.in_scale && i < Cin) ? d.in_scale[i] : 1.f; shs[i] = (d.in_scale && i < Cin) ? d.in_shift[i] : 0.f;
    }
    __syncthreads();
    const int ci4 = (Cin + 3) >> 2;
    const bool vout = (Cout % 4 == 0) && (d.y_ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(d.y) & 15u) == 0);
    float ssum[STATS ? CO : 1], ssq[STATS ? CO : 1];
    if (STATS) {
#pragma unroll
        for (int n = 0; n < CO; ++n) { ssum[n] = 0.f; ssq[n] = 0.f; }
    }
    for (long long m = (long long)blockIdx.x * kSkTile + tid; m < p.M; m += (long long)gridDim.x * kSkTile) {
        const float4* xr = reinterpret_cast<const float4*>(d.x + (size_t)m * d.x_ld);
        float4 xv[CI4];
#pragma unroll
        for (int j = 0; j < CI4; ++j) xv[j] = (j < ci4) ? __ldg(xr + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        float acc[CO];
#pragma unroll
        for (int n = 0; n < CO; ++n) acc[n] = bs[n];
#pragma unroll
        for (int j = 0; j < CI4; ++j) {
            if (j < ci4) {                                   // block-uniform
                const float xs[4] = {xv[j].x, xv[j].y, xv[j].z, xv[j].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int c = 4 * j + e;
                    float x1 = xs[e];
                    if (c >= Cin) x1 = 0.f;                  // row padding may hold anything
                    else if (d.in_scale) { x1 = fmaf(x1, scs[c], shs[c]); if (d.in_relu) x1 = fmaxf(x1, 0.f); }
                    const float4* wr = reinterpret_cast<const float4*>(Ws + c * CO);
#pragma unroll
                    for (int q4 = 0; q4 < CO / 4; ++q4) {
                        const float4 w4 = wr[q4];
                        acc[4 * q4] = fmaf(x1, w4.x, acc[4 * q4]); acc[4 * q4 + 1] = fmaf(x1, w4.y, acc[4 * q4 + 1]);
                        acc[4 * q4 + 2] = fmaf(x1, w4.z, acc[4 * q4 + 2]); acc[4 * q4 + 3] = fmaf(x1, w4.w, acc[4 * q4 + 3]);
                    }
                }
            }
        }
        const float rs = d.row_scale ? (d.row_scale[m] + d.row_scale_add) : 1.f;
        float* yr = d.y + (size_t)m * d.y_ld;
        if (STATS) {
#pragma unroll
            for (int n = 0; n < CO; ++n) { ssum[n] += acc[n]; ssq[n] = fmaf(acc[n], acc[n], ssq[n]); }
        }
#pragma unroll
        for (int n = 0; n < CO; ++n) acc[n] = apply_act(acc[n] * rs, d.act);
        if (vout) {
#pragma unroll
            for (int q4 = 0; q4 < CO / 4; ++q4) {
                if (4 * q4 < Cout) {
                    float4 o = make_float4(acc[4 * q4], acc[4 * q4 + 1], acc[4 * q4 + 2], acc[4 * q4 + 3]);
                    float4* dst = reinterpret_cast<float4*>(yr) + q4;
                    if (d.accumulate) { const float4 c4 = *dst; o.x += c4.x; o.y += c4.y; o.z += c4.z; o.w += c4.w; }
                    *dst = o;
                }
            }
        } else {
#pragma unroll
            for (int n = 0; n < CO; ++n)
                if (n < Cout) yr[n] = d.accumulate ? yr[n] + acc[n] : acc[n];
        }
    }
    if (STATS) {
#pragma unroll
        for (int n = 0; n < CO; ++n) {
            const float a = warp_sum(ssum[n]), b = warp_sum(ssq[n]);
            if (lane == 0) { red[warp][n] = a; red[warp][CO + n] = b; }
        }
        __syncthreads();
        for (int i = tid; i < 2 * CO; i += kSkTile) {
            const int n = i % CO;
            if (n < Cout) {
                float sacc = 0.f;
#pragma unroll
                for (int w = 0; w < kSkTile / 32; ++w) sacc += red[w][i];
                atomicAdd((i < CO ? d.stat_sum : d.stat_sumsq) + n, (double)sacc);
            }
        }
    }
}

template <int CO, int CI4>
static int launch_sk_row(const SkP& p, cudaStream_t st) {
    long long blocks = ((long long)p.M + kSkTile - 1) / kSkTile;
    const long long cap = (long long)kNumSMs * 8;
    if (blocks > cap) blocks = cap;
    if (p.d.stat_sum) skinny_fwd_row_kernel<CO, CI4, true><<<(int)blocks, kSkTile, 0, st>>>(p);
    else skinny_fwd_row_kernel<CO, CI4, false><<<(int)blocks, kSkTile, 0, st>>>(p);
    SAUNET_CHECK_LAUNCH("skinny_fwd_row_kernel");
    return SAUNET_OK;
}
template <int CO>
static int launch_sk_row_ci(const SkP& p, cudaStream_t st) {
    const int ci4 = (p.d.Cin + 3) / 4;
    if (ci4 <= 2) return launch_sk_row<CO, 2>(p, st);
    if (ci4 <= 5) return launch_sk_row<CO, 5>(p, st);
    return launch_sk_row<CO, 10>(p, st);
}

template <int CO>
static int launch_sk(const SkP& p, cudaStream_t st) {
    long long blocks = ((long long)p.M + kSkTile - 1) / kSkTile;
    const long long cap = (long long)kNumSMs * 8;
    if (blocks > cap) blocks = cap;
    skinny_fwd_kernel<CO><<<(int)blocks, kSkTile, 0, st>>>(p);
    SAUNET_CHECK_LAUNCH("skinny_fwd_kernel");
    return SAUNET_OK;
}

bool conv_skinny_eligible(const saunet_conv_desc* d) {
    if (d->KH != 1 || d->KW != 1 || d->sy != 1 || d->sx != 1 || d->offy != 0 || d->offx != 0) return false;
    if (d->osy != 1 || d->osx != 1 || d->oy0 != 0 || d->ox0 != 0) return false;
    if (d->Hg != d->Hin || d->Wg != d->Win || d->Hout != d->Hin || d->Wout != d->Win) return false;
    if (d->Cin > kSkMaxC || d->Cout > kSkMaxC) return false;
    return (long long)d->B * d->Hg * d->Wg >= 16384;        // tiny problems: any kernel will do
}

int conv_fwd_skinny(const saunet_conv_desc* d, cudaStream_t st) {
    SkP p; p.d = *d;
    const long long M = (long long)d->B * d->Hg * d->Wg;
    SAUNET_CHECK_ARG(M > 0 && M < (1ll << 31), SAUNET_ERR_BAD_SHAPE, "conv2d_fwd(skinny): bad M=%lld", M);
    p.M = (int)M;
    if (d->x_ld % 4 == 0 && aligned16(d->x) && !SAUNET_ENV_FLAG("SAUNET_SKINNY_OLD")) {      // 16-byte aligned rows: row-per-thread kernel
        if (d->Cout <= 4) return launch_sk_row_ci<4>(p, st);
        if (d->Cout <= 8) return launch_sk_row_ci<8>(p, st);
        if (d->Cout <= 16) return launch_sk_row_ci<16>(p, st);
        if (d->Cout <= 24) return launch_sk_row_ci<24>(p, st);
        if (d->Cout <= 32) return launch_sk_row_ci<32>(p, st);
        return launch_sk_row_ci<kSkMaxC>(p, st);
    }
    if (d->Cout <= 1) return launch_sk<1>(p, st);
    if (d->Cout <= 4) return launch_sk<4>(p, st);
    if (d->Cout <= 8) return launch_sk<8>(p, st);
    if (d->Cout <= 16) return launch_sk<16>(p, st);
    if (d->Cout <= 24) return launch_sk<24>(p, st);
    if (d->Cout <= 32) return launch_sk<32>(p, st);
    return launch_sk<kSkMaxC>(p, st);
}

// ---- weight gradient: dw[cb][ca] += sum_m pro(Q[m][cb]) * P[m][ca]   (1x1, Ca, Cb <= 40) ----
// Block = 256 threads; 64 pixels at a time are staged in shared memory; thread t < 100 owns a 4x4 tile of the
// 40x40 (padded) outer product and keeps it in registers over the block's whole pixel range.
constexpr int kSkPix = 64;

__global__ void __launch_bounds__(256) skinny_wgrad_kernel(const saunet_wgrad_desc d, long long M, long long pix_per_block) {
    __shared__ __align__(16) float Qs[kSkPix][kSkMaxC];
    __shared__ __align__(16) float Ps[kSkPix][kSkMaxC];
    const int tid = threadIdx.x;
    const int tb = tid / 10, ta = tid - tb * 10;             // 4x4 tile (cb4, ca4), valid for tid < 100
    const bool owner = tid < 100 && tb * 4 < d.Cb && ta * 4 < d.Ca;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const long long mbeg = (long long)blockIdx.x * pix_per_block;
    long long mend = mbeg + pix_per_block; if (mend > M) mend = M;
    // each thread stages 10 elements of Q and of P per 64-pixel chunk; the next chunk's loads are in flight while
    // the current chunk is multiplied
    constexpr int NL = kSkPix * kSkMaxC / 256;               // 10
    float qv[NL], pv[NL];
    auto fetch = [&](long long m0) {
#pragma unroll
        for (int k = 0; k < NL; ++k) {
            const int i = tid + 256 * k;
            const int px = i / kSkMaxC, c = i - px * kSkMaxC;
            const long long m = m0 + px;
            qv[k] = (m < mend && c < d.Cb) ? __ldg(d.q + (size_t)m * d.q_ld + c) : 0.f;
            pv[k] = (m < mend && c < d.Ca) ? __ldg(d.p + (size_t)m * d.p_ld + c) : 0.f;
        }
    };
    if (mbeg < mend) fetch(mbeg);
    for (long long m0 = mbeg; m0 < mend; m0 += kSkPix) {
        __syncthreads();
#pragma unroll
        for (int k = 0; k < NL; ++k) {
            const int i = tid + 256 * k;
            const int px = i / kSkMaxC, c = i - px * kSkMaxC;
            float q = qv[k];
            if (d.q_scale && c < d.Cb && m0 + px < mend) { q = fmaf(q, d.q_scale[c], d.q_shift[c]); if (d.q_relu) q = fmaxf(q, 0.f); }
            Qs[px][c] = q; Ps[px][c] = pv[k];
        }
        __syncthreads();
        if (m0 + kSkPix < mend) fetch(m0 + kSkPix);
        if (owner) {
#pragma unroll 8
            for (int px = 0; px < kSkPix; ++px) {
                const float4 q4 = *reinterpret_cast<const float4*>(&Qs[px][tb * 4]);
                const float4 p4 = *reinterpret_cast<const float4*>(&Ps[px][ta * 4]);
                const float qv[4] = {q4.x, q4.y, q4.z, q4.w}, pv[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(qv[i], pv[j], acc[i][j]);
            }
        }
    }
    if (owner) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int cb = tb * 4 + i, ca = ta * 4 + j;
                if (cb < d.Cb && ca < d.Ca) atomicAdd(d.dw + (size_t)cb * d.Ca + ca, acc[i][j]);
            }
    }
}



// ---- weight gradient, cp.async-pipelined pixel-split variant -------------------------------------------------------
// dw[cb][ca] += sum_m pro(Q[m][cb]) * P[m][ca].  The kernel above keeps a padded 40x40 outer product on 100 threads
// and exposes one global-load latency per 64-pixel chunk, so every call costs >= 0.15 ms whatever the channel counts.
// Here (a) chunks of 64 pixels stream through a 4-stage cp.async ring (16-byte copies when the rows are 16-byte
// aligned, 4-byte copies otherwise; rows past the range are zero-filled), (b) the NT = ceil(Cb/4)*ceil(Ca/4) register
// tiles are spread over all 256 threads: thread (tile t, pixel group g) accumulates pixels g, g+G, ... (G = 256/NT),
// applying the BN+ReLU prologue to its 4 Q channels on the fly, and (c) the G partial tiles are combined once per
// block through shared-memory atomics.  Work scales with Cb*Ca.
constexpr int kSwStages = 4;
constexpr int kSwPitch = kSkMaxC + 4;                      // 44 floats: rows 16-byte aligned

__device__ __forceinline__ void cp_async16(float* dst, const float* src, bool valid) {
    const unsigned n = valid ? 16u : 0u;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async4(float* dst, const float* src, bool valid) {
    const unsigned n = valid ? 4u : 0u;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src), "r"(n) : "memory");
}

template <bool ALIGNED>
__global__ void __launch_bounds__(256) skinny_wgrad_async_kernel(const saunet_wgrad_desc d, long long M, long long pix_per_block) {
    extern __shared__ __align__(16) float sw_smem[];
    float (*Qs)[kSkPix][kSwPitch] = reinterpret_cast<float (*)[kSkPix][kSwPitch]>(sw_smem);
    float (*Ps)[kSkPix][kSwPitch] = reinterpret_cast<float (*)[kSkPix][kSwPitch]>(sw_smem + kSwStages * kSkPix * kSwPitch);
    __shared__ float red[100 * 16];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cb4 = (d.Cb + 3) >> 2, ca4 = (d.Ca + 3) >> 2;
    const int NT = cb4 * ca4, G = 256 / NT;
    const int t = tid % NT, g = tid / NT;
    const bool worker = g < G;
    const int tb = t / ca4, ta = t - tb * ca4;
    for (int i = tid; i < NT * 16; i += 256) red[i] = 0.f;
    const long long mbeg = (long long)blockIdx.x * pix_per_block;
    long long mend = mbeg + pix_per_block; if (mend > M) mend = M;
    const int nchunk = mend > mbeg ? (int)((mend - mbeg + kSkPix - 1) / kSkPix) : 0;
    // stage a chunk: warp w copies rows w, w + 8, ...
    auto issue = [&](int ci) {
        if (ci < nchunk) {
            const int st = ci % kSwStages;
            const long long m0 = mbeg + (long long)ci * kSkPix;
#pragma unroll
            for (int k = 0; k < kSkPix / 8; ++k) {
                const int px = warp + 8 * k;
                const long long m = m0 + px;
                const bool rv = m < mend;
                const float* qrow = d.q + (size_t)(rv ? m : mbeg) * d.q_ld;
                const float* prow = d.p + (size_t)(rv ? m : mbeg) * d.p_ld;
                if (ALIGNED) {
                    if (lane < cb4) cp_async16(&Qs[st][px][lane * 4], qrow + lane * 4, rv);
                    else if (lane >= 16 && lane - 16 < ca4) cp_async16(&Ps[st][px][(lane - 16) * 4], prow + (lane - 16) * 4, rv);
                } else {
                    if (lane < d.Cb) cp_async4(&Qs[st][px][lane], qrow + lane, rv);
                    if (lane + 32 < d.Cb) cp_async4(&Qs[st][px][lane + 32], qrow + lane + 32, rv);
                    if (lane < d.Ca) cp_async4(&Ps[st][px][lane], prow + lane, rv);
                    if (lane + 32 < d.Ca) cp_async4(&Ps[st][px][lane + 32], prow + lane + 32, rv);
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // per-thread prologue constants of the tile's 4 Q channels; channels >= Cb contribute zero
    float qs[4], qh[4]; bool qe[4], pe[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int c = tb * 4 + e;
        qe[e] = worker && c < d.Cb; pe[e] = ta * 4 + e < d.Ca;
        qs[e] = (d.q_scale && qe[e]) ? d.q_scale[c] : 1.f; qh[e] = (d.q_scale && qe[e]) ? d.q_shift[c] : 0.f;
    }
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll
    for (int s0 = 0; s0 < kSwStages - 1; ++s0) issue(s0);
    for (int ci = 0; ci < nchunk; ++ci) {
        asm volatile("cp.async.wait_group %0;" ::"n"(kSwStages - 2) : "memory");
        __syncthreads();                                   // chunk ci landed for everyone; chunk ci-1 fully consumed
        issue(ci + kSwStages - 1);                         // refills the buffer of chunk ci-1
        if (worker) {
            const int st = ci % kSwStages;
            const long long m0 = mbeg + (long long)ci * kSkPix;
            int npx = (int)(mend - m0 < kSkPix ? mend - m0 : kSkPix);
            for (int px = g; px < npx; px += G) {
                const float4 q4 = *reinterpret_cast<const float4*>(&Qs[st][px][tb * 4]);
                const float4 p4 = *reinterpret_cast<const float4*>(&Ps[st][px][ta * 4]);
                float qa[4] = {q4.x, q4.y, q4.z, q4.w};
                float pa[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float q = qa[e];
                    if (d.q_scale) { q = fmaf(q, qs[e], qh[e]); if (d.q_relu) q = fmaxf(q, 0.f); }
                    qa[e] = qe[e] ? q : 0.f;                 // (row padding may hold anything)
                    pa[e] = pe[e] ? pa[e] : 0.f;
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(qa[i], pa[j], acc[i][j]);
            }
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (worker) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) atomicAdd(&red[t * 16 + i * 4 + j], acc[i][j]);
    }
    __syncthreads();
    for (int i = tid; i < NT * 16; i += 256) {
        const int tt = i >> 4, k = i & 15;
        const int cb = (tt / ca4) * 4 + (k >> 2), ca = (tt % ca4) * 4 + (k & 3);
        if (cb < d.Cb && ca < d.Ca) atomicAdd(d.dw + (size_t)cb * d.Ca + ca, red[i]);
    }
}

bool conv_wgrad_skinny_eligible(const saunet_wgrad_desc* d) {
    if (d->KH != 1 || d->KW != 1 || d->sy != 1 || d->sx != 1 || d->offy != 0 || d->offx != 0) return false;
    if (d->Hg != d->Hq || d->Wg != d->Wq) return false;
    if (d->Ca > kSkMaxC || d->Cb > kSkMaxC) return false;
    return (long long)d->B * d->Hg * d->Wg >= 16384;
}

int conv_wgrad_skinny(const saunet_wgrad_desc* d, cudaStream_t st) {
    const long long M = (long long)d->B * d->Hg * d->Wg;
    SAUNET_CHECK_ARG(M > 0, SAUNET_ERR_BAD_SHAPE, "conv2d_wgrad(skinny): empty problem");
    long long blocks = (long long)kNumSMs * 6;
    long long ppb = (M + blocks - 1) / blocks;
    ppb = (ppb + kSkPix - 1) / kSkPix * kSkPix;
    blocks = (M + ppb - 1) / ppb;
    if (!SAUNET_ENV_FLAG("SAUNET_SKINNY_OLD")) {
        const int smem = 2 * kSwStages * kSkPix * kSwPitch * 4;               // 90 KB: two blocks per SM
        static bool attr_set = false;
        if (!attr_set) {
            cudaFuncSetAttribute(skinny_wgrad_async_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            cudaFuncSetAttribute(skinny_wgrad_async_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            attr_set = true;
        }
        long long nb = (long long)kNumSMs * 2;
        long long pb = (M + nb - 1) / nb; pb = (pb + kSkPix - 1) / kSkPix * kSkPix;
        nb = (M + pb - 1) / pb;
        const bool al = d->q_ld % 4 == 0 && d->p_ld % 4 == 0 && aligned16(d->q) && aligned16(d->p);
        if (al) skinny_wgrad_async_kernel<true><<<(int)nb, 256, smem, st>>>(*d, M, pb);
        else skinny_wgrad_async_kernel<false><<<(int)nb, 256, smem, st>>>(*d, M, pb);
        SAUNET_CHECK_LAUNCH("skinny_wgrad_async_kernel");
        return SAUNET_OK;
    }
    skinny_wgrad_kernel<<<(int)blocks, 256, 0, st>>>(*d, M, ppb);
    SAUNET_CHECK_LAUNCH("skinny_wgrad_kernel");
    return SAUNET_OK;
}

}  // namespace saunet
