// dataprep.cu -- the integer / byte work either side of the network:
//   * saunet_argmax_u8: per-pixel class index of the logits (test_and_pack.py:119 / train.py:49:
//     `_, pred = torch.max(p1, dim=1)`; first maximum on ties) as uint8, for volume inference;
//   * saunet_edge_gt: the edge ground truth of the loader, data/ac17_dataloader.py:231-258 (mask_to_edges):
//     per class i in 1..3, dist = EDT(onehot_i) + EDT(1 - onehot_i) on the 1-padded map, kept where dist <= radius,
//     edge = any class kept.  EDT(onehot)+EDT(1-onehot) at p is the distance from p to the nearest pixel whose class-i
//     membership differs (outside the image = not a member), so with radius 2 the whole construct is a fixed stencil:
//     edge(p) = 1 iff some q with |p-q|^2 <= radius^2 has label(q) != label(p), out-of-image q counting as label 0.
//     (label(q) != label(p) <=> membership differs for class label(p) or label(q), one of which is >= 1.)
// Both are HBM-bound single passes: 1 B/pixel out, 8 B/pixel (int64 labels) or 4C B/pixel in.
#include "common.cuh"

namespace saunet {

__global__ void __launch_bounds__(256) argmax_u8_kernel(const float* __restrict__ logits, int ld, int C, long long npix, unsigned char* __restrict__ out) {
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < npix; p += (long long)gridDim.x * blockDim.x) {
        const float* z = logits + (size_t)p * ld;
        float best = __ldg(z); int bi = 0;
        for (int c = 1; c < C; ++c) { const float v = __ldg(z + c); if (v > best) { best = v; bi = c; } }
        out[p] = (unsigned char)bi;
    }
}

__global__ void __launch_bounds__(256) edge_gt_kernel(const long long* __restrict__ seg, int B, int H, int W, int radius, int num_classes,
                                                      float* __restrict__ out) {
    const long long n = (long long)B * H * W;
    const int r2 = radius * radius;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % W), y = (int)((i / W) % H);
        const long long base = i - (long long)y * W - x;                 // first pixel of the image
        long long l0 = seg[i];
        if (l0 < 1 || l0 > num_classes) l0 = 0;                          // classes 1..num_classes only (ac17_dataloader.py:239)
        bool e = false;
        for (int dy = -radius; dy <= radius && !e; ++dy)
            for (int dx = -radius; dx <= radius; ++dx) {
                if (dy * dy + dx * dx > r2 || (dy == 0 && dx == 0)) continue;
                const int yy = y + dy, xx = x + dx;
                long long l = 0;
                if (yy >= 0 && yy < H && xx >= 0 && xx < W) { l = seg[base + (long long)yy * W + xx]; if (l < 1 || l > num_classes) l = 0; }
                else if (yy < -1 || yy > H || xx < -1 || xx > W) continue;  // the reference pads by ONE pixel only
                if (l != l0) { e = true; break; }
            }
        out[i] = e ? 1.f : 0.f;
    }
}

}  // namespace saunet

using namespace saunet;

extern "C" int saunet_argmax_u8(const float* logits, int ld, int C, long long npix, unsigned char* out, void* stream) {
    SAUNET_CHECK_ARG(logits && out && C > 0 && C <= 255 && ld >= C && npix > 0, SAUNET_ERR_BAD_SHAPE, "argmax_u8: bad args");
    long long blocks = (npix + 255) / 256; if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    argmax_u8_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(logits, ld, C, npix, out);
    SAUNET_CHECK_LAUNCH("argmax_u8_kernel");
    return SAUNET_OK;
}

extern "C" int saunet_edge_gt(const long long* seg, int B, int H, int W, int radius, int num_classes, float* out, void* stream) {
    SAUNET_CHECK_ARG(seg && out && B > 0 && H > 0 && W > 0 && radius >= 1 && radius <= 8 && num_classes >= 1, SAUNET_ERR_BAD_SHAPE, "edge_gt: bad args");
    const long long n = (long long)B * H * W;
    long long blocks = (n + 255) / 256; if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    edge_gt_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(seg, B, H, W, radius, num_classes, out);
    SAUNET_CHECK_LAUNCH("edge_gt_kernel");
    return SAUNET_OK;
}
