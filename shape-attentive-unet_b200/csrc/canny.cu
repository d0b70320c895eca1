// canny.cu -- on-device replacement for the host Canny step of SAUNet.forward
// (models/models.py:358-364: np.mean(x, axis=1).astype(np.uint8); cv2.Canny(im, 10, 100)).
// Integer pipeline, bit-exact with OpenCV (apertureSize 3, L1 gradient): Sobel with replicated
// borders, |dx|+|dy| magnitude with a zero frame, 4-sector non-max suppression by the fixed-point
// tan(22.5) test, 8-connected hysteresis.  Removes the D2H copy + per-sample host loop + H2D copy.
#include "common.cuh"

namespace saunet {

// float mean over the channel axis (sequential fp32 sum / C), then the x86 float->uint8 cast
// (truncate to int32, keep the low byte) that numpy performs in the reference.
__global__ void __launch_bounds__(256) canny_u8_kernel(const float* __restrict__ x, int C, long long HW, long long n, uint8_t* __restrict__ im) {
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
        long long b = idx / HW, p = idx - b * HW;
        const float* src = x + (size_t)b * C * HW + p;
        float s = src[0];
        for (int c = 1; c < C; ++c) s = __fadd_rn(s, src[(size_t)c * HW]);
        float m = __fdiv_rn(s, (float)C);
        im[idx] = (uint8_t)(((int)m) & 0xFF);
    }
}

__device__ __forceinline__ int px(const uint8_t* im, int H, int W, int y, int x) {
    y = min(max(y, 0), H - 1); x = min(max(x, 0), W - 1);
    return (int)im[(size_t)y * W + x];
}
__device__ __forceinline__ void sobel(const uint8_t* im, int H, int W, int y, int x, int& gx, int& gy) {
    int a = px(im, H, W, y - 1, x - 1), b = px(im, H, W, y - 1, x), c = px(im, H, W, y - 1, x + 1);
    int d = px(im, H, W, y, x - 1), f = px(im, H, W, y, x + 1);
    int g = px(im, H, W, y + 1, x - 1), h = px(im, H, W, y + 1, x), i = px(im, H, W, y + 1, x + 1);
    gx = (c + 2 * f + i) - (a + 2 * d + g);
    gy = (g + 2 * h + i) - (a + 2 * b + c);
}
__device__ __forceinline__ int mag_at(const uint8_t* im, int H, int W, int y, int x) {
    if (y < 0 || y >= H || x < 0 || x >= W) return 0;      // zero frame around the magnitude image
    int gx, gy; sobel(im, H, W, y, x, gx, gy);
    return abs(gx) + abs(gy);
}

// map: 0 = weak candidate, 1 = not an edge, 2 = edge
__global__ void __launch_bounds__(256) canny_nms_kernel(const uint8_t* __restrict__ im, int B, int H, int W, int low, int high, uint8_t* __restrict__ map) {
    const long long n = (long long)B * H * W;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
        int x = (int)(idx % W); long long q = idx / W; int y = (int)(q % H); int b = (int)(q / H);
        const uint8_t* I = im + (size_t)b * H * W;
        int xs, ys; sobel(I, H, W, y, x, xs, ys);
        const int m = abs(xs) + abs(ys);
        uint8_t r = 1;
        if (m > low) {
            const int ax = abs(xs), ay = abs(ys) << 15;
            const int tg22x = ax * 13573;
            bool keep;
            if (ay < tg22x) keep = (m > mag_at(I, H, W, y, x - 1)) && (m >= mag_at(I, H, W, y, x + 1));
            else {
                const int tg67x = tg22x + (ax << 16);
                if (ay > tg67x) keep = (m > mag_at(I, H, W, y - 1, x)) && (m >= mag_at(I, H, W, y + 1, x));
                else {
                    const int s = ((xs ^ ys) < 0) ? -1 : 1;
                    keep = (m > mag_at(I, H, W, y - 1, x - s)) && (m > mag_at(I, H, W, y + 1, x + s));
                }
            }
            if (keep) r = (m > high) ? 2 : 0;
        }
        map[idx] = r;
    }
}

// one CTA per image: iterate "weak pixel with an edge neighbour becomes an edge" to the fixed point
// (the unique result of OpenCV's stack-based flood fill), then emit {0,255} floats.
__global__ void __launch_bounds__(1024) canny_hysteresis_kernel(uint8_t* map_all, int H, int W, float* __restrict__ out) {
    volatile uint8_t* map = map_all + (size_t)blockIdx.x * H * W;
    float* o = out + (size_t)blockIdx.x * H * W;
    const int n = H * W;
    // each thread owns a contiguous run of pixels so edges propagate along a run within one sweep
    const int per = (n + blockDim.x - 1) / blockDim.x;
    const int beg = threadIdx.x * per, end = min(n, beg + per);
    int changed = 1;
    while (changed) {
        int ch = 0;
        for (int dir = 0; dir < 2; ++dir) {
            for (int k = 0; k < end - beg; ++k) {
                const int i = dir ? (end - 1 - k) : (beg + k);
                if (map[i] != 0) continue;
                const int y = i / W, x = i - y * W;
                bool hit = false;
#pragma unroll
                for (int dy = -1; dy <= 1; ++dy) {
                    const int yy = y + dy;
                    if (yy < 0 || yy >= H) continue;
#pragma unroll
                    for (int dx = -1; dx <= 1; ++dx) {
                        const int xx = x + dx;
                        if ((dy | dx) == 0 || xx < 0 || xx >= W) continue;
                        hit |= (map[yy * W + xx] == 2);
                    }
                }
                if (hit) { map[i] = 2; ch = 1; }
            }
        }
        changed = __syncthreads_or(ch);
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) o[i] = (map[i] == 2) ? 255.f : 0.f;
}

}  // namespace saunet

using namespace saunet;

extern "C" long long saunet_canny_workspace_bytes(int B, int H, int W) {
    if (B <= 0 || H <= 0 || W <= 0) return 0;
    return 2ll * B * H * W + 256;
}

extern "C" int saunet_canny_fwd(const float* x_nchw, int B, int C, int H, int W, int low, int high, float* out, void* workspace,
                                long long workspace_bytes, void* stream) {
    SAUNET_CHECK_ARG(x_nchw && out && B > 0 && C > 0 && H > 0 && W > 0, SAUNET_ERR_BAD_SHAPE, "canny_fwd: bad args");
    SAUNET_CHECK_ARG((long long)H * W < (1ll << 30), SAUNET_ERR_BAD_SHAPE, "canny_fwd: image too large");
    SAUNET_CHECK_ARG(workspace && workspace_bytes >= saunet_canny_workspace_bytes(B, H, W), SAUNET_ERR_WORKSPACE,
                     "canny_fwd: workspace too small (%lld < %lld)", workspace_bytes, saunet_canny_workspace_bytes(B, H, W));
    cudaStream_t st = (cudaStream_t)stream;
    const long long n = (long long)B * H * W;
    uint8_t* im = (uint8_t*)workspace;
    uint8_t* map = im + ((n + 127) / 128) * 128;
    int blocks = (int)((n + 255) / 256); if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    canny_u8_kernel<<<blocks, 256, 0, st>>>(x_nchw, C, (long long)H * W, n, im);
    SAUNET_CHECK_LAUNCH("canny_u8_kernel");
    canny_nms_kernel<<<blocks, 256, 0, st>>>(im, B, H, W, low, high, map);
    SAUNET_CHECK_LAUNCH("canny_nms_kernel");
    canny_hysteresis_kernel<<<B, 1024, 0, st>>>(map, H, W, out);
    SAUNET_CHECK_LAUNCH("canny_hysteresis_kernel");
    return SAUNET_OK;
}
