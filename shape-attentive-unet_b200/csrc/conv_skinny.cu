// conv_skinny.cu -- 1x1 convolutions with very few channels on very many pixels (the gated shape stream at full
// resolution: C+1 -> C+1 and C+1 -> 1 gate convs, fuse 8->1, cw 2->1, expand 1->32, final 32->4 and their data /
// weight gradients; models/GSConv.py:38-45, models/models.py:291-301,324).  These are HBM-bound (a few hundred bytes
// per pixel, <= 40x40 MACs): one thread per pixel, weights broadcast from shared memory, no tensor cores, any
// channel count / alignment.  Same descriptor and the same fused prologue / epilogue as the GEMM kernels.
#include "common.cuh"

namespace saunet {

constexpr int kSkMaxC = 40;

struct SkP {
    saunet_conv_desc d;
    int M;
};

// ---- forward / data gradient: y[m][0:Cout) = act(rs[m] * (sum_c pro(x[m][c]) * W[c][n] + bias[n])) ----
template <int CO>
__global__ void __launch_bounds__(256) skinny_fwd_kernel(const SkP p) {
    __shared__ __align__(16) float Ws[kSkMaxC * CO];
    __shared__ float bs[CO], scs[kSkMaxC], shs[kSkMaxC];
    __shared__ float red[8][2 * CO];
    const saunet_conv_desc& d = p.d;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < d.Cin * CO; i += 256) {
        const int c = i / CO, n = i - c * CO;
        Ws[i] = n < d.Cout ? d.w[(size_t)c * d.Cout + n] : 0.f;
    }
    for (int i = tid; i < CO; i += 256) bs[i] = (d.bias && i < d.Cout) ? d.bias[i] : 0.f;
    for (int i = tid; i < d.Cin; i += 256) { scs[i] = d.in_scale ? d.in_scale[i] : 1.f; shs[i] = d.in_scale ? d.in_shift[i] : 0.f; }
    __syncthreads();
    float ssum[CO], ssq[CO];
#pragma unroll
    for (int n = 0; n < CO; ++n) { ssum[n] = 0.f; ssq[n] = 0.f; }
    for (int m = blockIdx.x * 256 + tid; m < p.M; m += gridDim.x * 256) {
        const float* xp = d.x + (size_t)m * d.x_ld;
        float acc[CO];
#pragma unroll
        for (int n = 0; n < CO; ++n) acc[n] = bs[n];
        for (int c = 0; c < d.Cin; ++c) {
            float xv = __ldg(xp + c);
            if (d.in_scale) { xv = fmaf(xv, scs[c], shs[c]); if (d.in_relu) xv = fmaxf(xv, 0.f); }
            const float* wr = Ws + c * CO;
#pragma unroll
            for (int n = 0; n < CO; ++n) acc[n] = fmaf(xv, wr[n], acc[n]);
        }
        const float rs = d.row_scale ? (d.row_scale[m] + d.row_scale_add) : 1.f;
        float* yp = d.y + (size_t)m * d.y_ld;
#pragma unroll
        for (int n = 0; n < CO; ++n) {
            if (n < d.Cout) {
                ssum[n] += acc[n]; ssq[n] = fmaf(acc[n], acc[n], ssq[n]);
                const float o = apply_act(acc[n] * rs, d.act);
                yp[n] = d.accumulate ? yp[n] + o : o;
            }
        }
    }
    if (d.stat_sum) {
#pragma unroll
        for (int n = 0; n < CO; ++n) {
            const float a = warp_sum(ssum[n]), b = warp_sum(ssq[n]);
            if (lane == 0) { red[warp][n] = a; red[warp][CO + n] = b; }
        }
        __syncthreads();
        for (int i = tid; i < 2 * CO; i += 256) {
            const int n = i % CO;
            if (n < d.Cout) {
                float s = 0.f;
#pragma unroll
                for (int w = 0; w < 8; ++w) s += red[w][i];
                atomicAdd((i < CO ? d.stat_sum : d.stat_sumsq) + n, (double)s);
            }
        }
    }
}

template <int CO>
static int launch_sk(const SkP& p, cudaStream_t st) {
    long long blocks = ((long long)p.M + 255) / 256;
    const long long cap = (long long)kNumSMs * 8;
    if (blocks > cap) blocks = cap;
    skinny_fwd_kernel<CO><<<(int)blocks, 256, 0, st>>>(p);
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
    for (long long m0 = mbeg; m0 < mend; m0 += kSkPix) {
        __syncthreads();
        for (int i = tid; i < kSkPix * kSkMaxC; i += 256) {
            const int px = i / kSkMaxC, c = i - px * kSkMaxC;
            const long long m = m0 + px;
            float q = 0.f, pv = 0.f;
            if (m < mend) {
                if (c < d.Cb) {
                    q = __ldg(d.q + (size_t)m * d.q_ld + c);
                    if (d.q_scale) { q = fmaf(q, d.q_scale[c], d.q_shift[c]); if (d.q_relu) q = fmaxf(q, 0.f); }
                }
                if (c < d.Ca) pv = __ldg(d.p + (size_t)m * d.p_ld + c);
            }
            Qs[px][c] = q; Ps[px][c] = pv;
        }
        __syncthreads();
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
    skinny_wgrad_kernel<<<(int)blocks, 256, 0, st>>>(*d, M, ppb);
    SAUNET_CHECK_LAUNCH("skinny_wgrad_kernel");
    return SAUNET_OK;
}

}  // namespace saunet
