// loss.cu -- fused DualLoss (Dice + weighted CE + edge BCE) forward / backward.
// Restates loss.py:51-88,149-159 of the reference as two single-pass HBM-bound kernels over the
// [npix][C] logits: fwd accumulates the 2+2C+1 batch sums (fp64 atomics, one per CTA), bwd recomputes
// the softmax and emits d(logits) and d(edge) in one pass.
// The forward pass also counts the training-branch metrics of SegmentationModule (models/models.py:51-74,92:
// pixel accuracy + per-class Jaccard of round(softmax(logits))) from the softmax it already holds, so the metric
// costs no extra pass over the logits and no torch kernels.
#include "common.cuh"

namespace saunet {

constexpr int kMaxC = 8;
constexpr float kDiceEps = 1e-7f;

template <int NQ>
__device__ __forceinline__ void block_reduce_atomic(float (&v)[NQ], double* out, int nq) {
    __shared__ float red[NQ][8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NQ; ++i) { float s = warp_sum(v[i]); if (lane == 0) red[i][w] = s; }
    __syncthreads();
    if (threadIdx.x < nq) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) s += red[threadIdx.x][k];
        atomicAdd(out + threadIdx.x, (double)s);
    }
}

__device__ __forceinline__ void softmax_c(const float* z, int C, float* prob, float* logp) {
    float m = z[0];
    for (int c = 1; c < C; ++c) m = fmaxf(m, z[c]);
    float s = 0.f;
    for (int c = 0; c < C; ++c) s += expf(z[c] - m);
    float ls = logf(s);
    for (int c = 0; c < C; ++c) { logp[c] = z[c] - m - ls; prob[c] = expf(logp[c]); }
}

__global__ void __launch_bounds__(256) dual_loss_fwd_kernel(const float* __restrict__ logits, int l_ld, const float* __restrict__ edge,
                                                            const long long* __restrict__ seg_t, const float* __restrict__ edge_t,
                                                            long long npix, int C, const float* __restrict__ cw, double* acc,
                                                            int* __restrict__ counts) {
    float q[2 + 2 * kMaxC + 1];
#pragma unroll
    for (int i = 0; i < 2 + 2 * kMaxC + 1; ++i) q[i] = 0.f;
    // metric counters: [0] correct & labelled, [1] labelled (label >= 1), then per class i>=1: [2+3(i-1)] = |label==i & pred==i|,
    // [+1] = |label==i|, [+2] = |pred==i|; last: labels outside [0,C) (skipped by every sum)
    int cnt[2 + 3 * (kMaxC - 1) + 1];
#pragma unroll
    for (int i = 0; i < 2 + 3 * (kMaxC - 1) + 1; ++i) cnt[i] = 0;
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < npix; p += (long long)gridDim.x * blockDim.x) {
        float z[kMaxC], pr[kMaxC], lp[kMaxC];
        for (int c = 0; c < C; ++c) z[c] = __ldg(logits + (size_t)p * l_ld + c);
        softmax_c(z, C, pr, lp);
        const long long tl = seg_t[p];
        const bool tval = tl >= 0 && tl < C;         // an out-of-range label (255, -100 ...) never indexes anything
        const int t = tval ? (int)tl : -1;
        const float w = tval ? (cw ? cw[t] : 1.f) : 0.f;
        for (int c = 0; c < C; ++c) {
            const float oh = (c == t) ? 1.f : 0.f;
            if (c == t) { q[0] -= w * lp[c]; q[1] += w; }
            q[2 + c] += pr[c] * oh;
            q[2 + kMaxC + c] += pr[c] + oh;
        }
        if (counts) {
            // torch.max(round(softmax(z)).long(), dim=1): the first class whose probability rounds to 1 (p > 0.5; exactly
            // 0.5 rounds to even = 0), class 0 when none does
            int pred = 0;
            for (int c = C - 1; c >= 1; --c) if (rintf(pr[c]) == 1.f) pred = c;
            if (rintf(pr[0]) == 1.f) pred = 0;
            if (!tval) cnt[2 + 3 * (kMaxC - 1)]++;
            if (t >= 1) { cnt[1]++; if (pred == t) cnt[0]++; }
#pragma unroll
            for (int i = 1; i < kMaxC; ++i) {
                if (i < C) {
                    const bool v = t == i, pp = pred == i;
                    cnt[2 + 3 * (i - 1)] += (v && pp); cnt[3 + 3 * (i - 1)] += v; cnt[4 + 3 * (i - 1)] += pp;
                }
            }
        }
        if (edge) {
            const float pe = __ldg(edge + p), te = __ldg(edge_t + p);
            const float l1 = fmaxf(logf(pe), -100.f), l0 = fmaxf(log1pf(-pe), -100.f);
            q[2 + 2 * kMaxC] -= te * l1 + (1.f - te) * l0;
        }
    }
    // compact to the public layout [0]=sum w*nll, [1]=sum w, [2..2+C)=I, [2+C..2+2C)=Card, [2+2C]=bce
    float v[2 + 2 * kMaxC + 1];
    v[0] = q[0]; v[1] = q[1];
    for (int c = 0; c < kMaxC; ++c) { v[2 + c] = 0.f; v[2 + kMaxC + c] = 0.f; }
    for (int c = 0; c < C; ++c) { v[2 + c] = q[2 + c]; v[2 + C + c] = q[2 + kMaxC + c]; }
    v[2 + 2 * C] = q[2 + 2 * kMaxC];
    block_reduce_atomic<2 + 2 * kMaxC + 1>(v, acc, 2 + 2 * C + 1);
    if (counts) {
        __shared__ int cred[2 + 3 * (kMaxC - 1) + 1];
        if (threadIdx.x < 2 + 3 * (kMaxC - 1) + 1) cred[threadIdx.x] = 0;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 2 + 3 * (kMaxC - 1) + 1; ++i) {
            int s = cnt[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if ((threadIdx.x & 31) == 0 && s) atomicAdd(&cred[i], s);
        }
        __syncthreads();
        if (threadIdx.x < 2 + 3 * (kMaxC - 1) + 1 && cred[threadIdx.x]) atomicAdd(counts + threadIdx.x, cred[threadIdx.x]);
    }
}

__global__ void dual_loss_finalize_kernel(const double* __restrict__ acc, long long npix, int C, int has_edge, int parts, float* loss,
                                          const int* __restrict__ counts) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (counts) {
        // models/models.py:57-72: acc = correct / (labelled + 1e-10); jaccard_i = anb / (|v| + |p| - anb + 1e-10), 0 if > 1
        loss[4] = (float)counts[0] / ((float)counts[1] + 1e-10f);
        for (int i = 1; i < C; ++i) {
            const float anb = (float)counts[2 + 3 * (i - 1)];
            const float j = anb / ((float)counts[3 + 3 * (i - 1)] + (float)counts[4 + 3 * (i - 1)] - anb + 1e-10f);
            loss[4 + i] = j <= 1.f ? j : 0.f;
        }
    }
    double ce = acc[0] / acc[1];
    double d = 0.0;
    for (int c = 0; c < C; ++c) d += 2.0 * acc[2 + c] / (acc[2 + C + c] + (double)kDiceEps);
    double dice = 1.0 - d / C;
    double bce = has_edge ? acc[2 + 2 * C] / (double)npix : 0.0;
    double total = 0.0;
    if (parts & 1) total += dice;
    if (parts & 2) total += ce;
    if (parts & 4) total += bce;
    loss[0] = (float)total; loss[1] = (float)dice; loss[2] = (float)ce; loss[3] = (float)bce;
}

__global__ void __launch_bounds__(256) dual_loss_bwd_kernel(const float* __restrict__ logits, int l_ld, const float* __restrict__ edge,
                                                            const long long* __restrict__ seg_t, const float* __restrict__ edge_t,
                                                            long long npix, int C, const float* __restrict__ cw, const double* __restrict__ acc,
                                                            const float* __restrict__ dloss, float* __restrict__ dlogits, int dl_ld,
                                                            float* __restrict__ dedge, int parts) {
    const float g = dloss ? dloss[0] : 1.f;
    const float inv_w = (parts & 2) ? (float)(1.0 / acc[1]) : 0.f;
    const float dsel = (parts & 1) ? 1.f : 0.f;
    float k1[kMaxC], k2[kMaxC];     // dice: d/dprob_c = -(1/C) * ( 2*oh/(Card+eps) - 2*I/(Card+eps)^2 )
    for (int c = 0; c < C; ++c) {
        double den = acc[2 + C + c] + (double)kDiceEps;
        k1[c] = dsel * (float)(-2.0 / (den * C));
        k2[c] = dsel * (float)(2.0 * acc[2 + c] / (den * den * C));
    }
    const float inv_n = 1.f / (float)npix;
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < npix; p += (long long)gridDim.x * blockDim.x) {
        float z[kMaxC], pr[kMaxC], lp[kMaxC], a[kMaxC];
        for (int c = 0; c < C; ++c) z[c] = __ldg(logits + (size_t)p * l_ld + c);
        softmax_c(z, C, pr, lp);
        const long long tl = seg_t[p];
        const int t = (tl >= 0 && tl < C) ? (int)tl : -1;
        const float w = t >= 0 ? (cw ? cw[t] : 1.f) * inv_w : 0.f;
        float dot = 0.f;
        for (int c = 0; c < C; ++c) { a[c] = (c == t ? k1[c] : 0.f) + k2[c]; dot = fmaf(a[c], pr[c], dot); }
        for (int c = 0; c < C; ++c) {
            float dz = w * (pr[c] - (c == t ? 1.f : 0.f)) + pr[c] * (a[c] - dot);
            dlogits[(size_t)p * dl_ld + c] = g * dz;
        }
        if (edge) {
            const float pe = __ldg(edge + p), te = __ldg(edge_t + p);
            dedge[p] = (parts & 4) ? g * inv_n * (pe - te) / fmaxf((1.f - pe) * pe, 1e-12f) : 0.f;
        }
    }
}

}  // namespace saunet

using namespace saunet;

extern "C" int saunet_dual_loss_fwd(const float* logits, int l_ld, const float* edge, const long long* seg_t, const float* edge_t,
                                    long long npix, int C, const float* class_w, int parts, double* acc, float* loss,
                                    int* counts, void* stream) {
    SAUNET_CHECK_ARG(logits && seg_t && acc && loss && npix > 0, SAUNET_ERR_BAD_SHAPE, "dual_loss_fwd: bad args");
    SAUNET_CHECK_ARG(C >= 2 && C <= kMaxC && l_ld >= C, SAUNET_ERR_BAD_SHAPE, "dual_loss_fwd: C=%d unsupported (2..8)", C);
    SAUNET_CHECK_ARG((edge == nullptr) == (edge_t == nullptr), SAUNET_ERR_BAD_SHAPE, "dual_loss_fwd: edge/edge_t mismatch");
    long long blocks = (npix + 255) / 256; if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    dual_loss_fwd_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(logits, l_ld, edge, seg_t, edge_t, npix, C, class_w, acc, counts);
    SAUNET_CHECK_LAUNCH("dual_loss_fwd_kernel");
    dual_loss_finalize_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(acc, npix, C, edge != nullptr, parts, loss, counts);
    SAUNET_CHECK_LAUNCH("dual_loss_finalize_kernel");
    return SAUNET_OK;
}

extern "C" int saunet_dual_loss_bwd(const float* logits, int l_ld, const float* edge, const long long* seg_t, const float* edge_t,
                                    long long npix, int C, const float* class_w, const double* acc, const float* dloss,
                                    float* dlogits, int dl_ld, float* dedge, int parts, void* stream) {
    SAUNET_CHECK_ARG(logits && seg_t && acc && dlogits && npix > 0, SAUNET_ERR_BAD_SHAPE, "dual_loss_bwd: bad args");
    SAUNET_CHECK_ARG(C >= 2 && C <= kMaxC && l_ld >= C && dl_ld >= C, SAUNET_ERR_BAD_SHAPE, "dual_loss_bwd: C=%d unsupported (2..8)", C);
    SAUNET_CHECK_ARG(!edge || (edge_t && dedge), SAUNET_ERR_BAD_SHAPE, "dual_loss_bwd: edge given without edge_t/dedge");
    long long blocks = (npix + 255) / 256; if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    dual_loss_bwd_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(logits, l_ld, edge, seg_t, edge_t, npix, C, class_w, acc, dloss, dlogits, dl_ld, dedge, parts);
    SAUNET_CHECK_LAUNCH("dual_loss_bwd_kernel");
    return SAUNET_OK;
}
