// optim.cu -- fused multi-tensor optimizer step over the flat parameter / gradient arenas.
//
// Replaces the reference's per-parameter Python loops (radam.py:15-78 RAdam; torch.optim.SGD / Adam as built by
// train.py:188-207 from the two parameter groups of train.py:166-185: conv / linear weights with weight decay,
// biases and BatchNorm affine parameters without) by ONE HBM-bound pass: every parameter of the model lives in one
// flat fp32 buffer (saunet_b200.parallel.GradArena.flatten_params) next to its gradient, so a step is
// w, g, m, v read once and w, m, v written once (28 B per parameter), whatever the number of tensors.
// Per-tensor hyper-parameters (learning rate, weight decay of the tensor's group) come from a small device table
// indexed by a binary search on the element offset; the step counter lives on the device and the bias corrections /
// RAdam rectification are computed from it in the kernel, so a step is CUDA-graph capturable and needs no host value.
#include "common.cuh"

namespace saunet {

struct OptSeg { long long begin; float lr, wd; };      // elements [begin, next.begin): one parameter tensor

__device__ __forceinline__ int seg_of(const OptSeg* __restrict__ seg, int nseg, long long i) {
    int lo = 0, hi = nseg - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (seg[mid].begin <= i) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// kind 0: SGD (torch.optim.SGD, nesterov=False, dampening 0): d = g + wd*w; buf = first ? d : mu*buf + d; w -= lr*buf
// kind 1: Adam (torch.optim.Adam): d = g + wd*w; m,v EMA; w -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
// kind 2: RAdam (radam.py:15-78): v,m EMA of g; w -= wd*lr*w; then w -= step_size * m / (sqrt(v) + eps) if N_sma >= 5
//         else w -= step_size * m, step_size / N_sma from the step count exactly as radam.py:52-65
__global__ void __launch_bounds__(256) optimizer_step_kernel(int kind, float* __restrict__ w, const float* __restrict__ g,
                                                             float* __restrict__ m, float* __restrict__ v, long long n4,
                                                             const OptSeg* __restrict__ seg, int nseg, const int* __restrict__ step_ptr,
                                                             float beta1, float beta2, float eps, float momentum) {
    const int t = *step_ptr + 1;                       // this step's index (1-based); the counter is bumped by a second tiny kernel
    // step-dependent scalars in double, as the Python reference computes them
    double bc1 = 1.0, bc2 = 1.0, nsma = 0.0, rect = 1.0;
    if (kind >= 1) {
        const double b1t = pow((double)beta1, (double)t), b2t = pow((double)beta2, (double)t);
        bc1 = 1.0 - b1t; bc2 = 1.0 - b2t;
        if (kind == 2) {
            const double nmax = 2.0 / (1.0 - (double)beta2) - 1.0;
            nsma = nmax - 2.0 * t * b2t / (1.0 - b2t);
            if (nsma >= 5.0) rect = sqrt((1.0 - b2t) * (nsma - 4.0) / (nmax - 4.0) * (nsma - 2.0) / nsma * nmax / (nmax - 2.0));
        }
    }
    for (long long i4 = blockIdx.x * (long long)blockDim.x + threadIdx.x; i4 < n4; i4 += (long long)gridDim.x * blockDim.x) {
        const long long i = i4 * 4;
        const OptSeg sg = seg[seg_of(seg, nseg, i)];
        float4 wv = *reinterpret_cast<const float4*>(w + i);
        const float4 gv = __ldg(reinterpret_cast<const float4*>(g + i));
        float4 mv = *reinterpret_cast<const float4*>(m + i);
        float ww[4] = {wv.x, wv.y, wv.z, wv.w}, gg[4] = {gv.x, gv.y, gv.z, gv.w}, mm[4] = {mv.x, mv.y, mv.z, mv.w};
        if (kind == 0) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float d = fmaf(sg.wd, ww[e], gg[e]);
                mm[e] = (momentum != 0.f) ? (t == 1 ? d : fmaf(momentum, mm[e], d)) : d;
                ww[e] = fmaf(-sg.lr, mm[e], ww[e]);
            }
            if (momentum != 0.f) *reinterpret_cast<float4*>(m + i) = make_float4(mm[0], mm[1], mm[2], mm[3]);
        } else {
            float4 vv4 = *reinterpret_cast<const float4*>(v + i);
            float vv[4] = {vv4.x, vv4.y, vv4.z, vv4.w};
            if (kind == 1) {
                const float step = (float)((double)sg.lr / bc1), rs2 = (float)(1.0 / sqrt(bc2));
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float d = fmaf(sg.wd, ww[e], gg[e]);
                    mm[e] = fmaf(beta1, mm[e], (1.f - beta1) * d);
                    vv[e] = fmaf(beta2, vv[e], (1.f - beta2) * d * d);
                    ww[e] -= step * mm[e] / (sqrtf(vv[e]) * rs2 + eps);
                }
            } else {
                const float step = (float)((double)sg.lr * rect / bc1);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    vv[e] = fmaf(beta2, vv[e], (1.f - beta2) * gg[e] * gg[e]);
                    mm[e] = fmaf(beta1, mm[e], (1.f - beta1) * gg[e]);
                    if (sg.wd != 0.f) ww[e] = fmaf(-sg.wd * sg.lr, ww[e], ww[e]);
                    ww[e] -= (nsma >= 5.0) ? step * mm[e] / (sqrtf(vv[e]) + eps) : step * mm[e];
                }
            }
            *reinterpret_cast<float4*>(m + i) = make_float4(mm[0], mm[1], mm[2], mm[3]);
            *reinterpret_cast<float4*>(v + i) = make_float4(vv[0], vv[1], vv[2], vv[3]);
        }
        *reinterpret_cast<float4*>(w + i) = make_float4(ww[0], ww[1], ww[2], ww[3]);
    }
}

__global__ void optimizer_bump_kernel(int* step_ptr) { if (threadIdx.x == 0 && blockIdx.x == 0) ++*step_ptr; }

}  // namespace saunet

using namespace saunet;

extern "C" int saunet_optimizer_step(int kind, float* w, const float* g, float* m, float* v, long long n, const void* seg_table,
                                     int nseg, int* step_counter, float beta1, float beta2, float eps, float momentum, void* stream) {
    SAUNET_CHECK_ARG(kind >= 0 && kind <= 2, SAUNET_ERR_BAD_SHAPE, "optimizer_step: kind must be 0 (sgd), 1 (adam) or 2 (radam)");
    SAUNET_CHECK_ARG(w && g && seg_table && step_counter && n > 0 && nseg > 0 && n % 4 == 0, SAUNET_ERR_BAD_SHAPE, "optimizer_step: bad args");
    SAUNET_CHECK_ARG((kind == 0 && (m || momentum == 0.f)) || (kind > 0 && m && v), SAUNET_ERR_BAD_SHAPE, "optimizer_step: state buffers missing");
    SAUNET_CHECK_ARG(aligned16(w) && aligned16(g) && (!m || aligned16(m)) && (!v || aligned16(v)), SAUNET_ERR_BAD_ALIGN, "optimizer_step: 16-byte alignment");
    const long long n4 = n / 4;
    long long blocks = (n4 + 255) / 256; if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    optimizer_step_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(kind, w, g, m, v, n4, static_cast<const OptSeg*>(seg_table), nseg,
                                                                        step_counter, beta1, beta2, eps, momentum);
    SAUNET_CHECK_LAUNCH("optimizer_step_kernel");
    optimizer_bump_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(step_counter);
    SAUNET_CHECK_LAUNCH("optimizer_bump_kernel");
    return SAUNET_OK;
}
