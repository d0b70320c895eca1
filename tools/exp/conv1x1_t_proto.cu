// PROTOTYPE of csrc/conv_pw_t.cu (standalone, not linked into libsaunet_b200.so; measured on B200: 1.8e-6 error,
// 0.083 ms = 181 TFLOP/s at M 262144, K 224): pointwise (1x1) convolution with the OUTPUT CHANNELS on the TMEM lanes and the PIXELS as the MMA N dimension,
//
//     Y^T[co (128 lanes)][pixel (256 columns)] += W[co][k] * X[pixel][k]^T        (3xTF32)
//
// i.e. A = the pre-tiled weight image (exactly the [k_block][hi,lo][128][32] image conv_tc.cu already streams as its
// B operand), B = the 256-pixel activation tile the producers write (K-major rows, as today).  Compared with
// conv_tc.cu (pixels on the lanes, N = 128 channels):
//   * each MMA is M=128 x N=256 x K=8: 12 KB of operand reads per 132 math clocks instead of 8 KB per 66 -- 25 % less
//     shared-memory operand traffic per FLOP (the 1x1 layers are bound by shared-memory bandwidth: MMA operand reads
//     + producer stores + weight copies ~ 200 B/clk/SM against ~128 available),
//   * the epilogue needs no transposition: lane = channel, so for every pixel the 32 lanes of a warp store 32
//     consecutive floats (one 128-byte line), and the BatchNorm statistics are plain per-thread sums over the
//     thread's columns (no shuffles, no shared-memory reduction),
//   * half as many MMA instructions / barrier round trips per pixel.
// Standalone test: Cout = 128, K % 32 == 0, M % 256 == 0, no prologue / bias; checks 512 rows against fp64 and times
// the whole problem.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../shape-attentive-unet_b200/csrc
//                           -I../../include -o conv1x1_t_proto conv1x1_t_proto.cu && ./conv1x1_t_proto [M] [K]
#include "tc_common.cuh"
#include <vector>
#include <cmath>
#include <cstdlib>
using namespace saunet;
namespace saunet { void set_error(const char*, ...) {} void count_launch(int) {} void note_kernel(const char*) {} }

constexpr int kNP = 256;                 // pixels per tile (MMA N)
constexpr int kProd = 512, kEpi = 256, kThreads = kProd + kEpi + 64;
constexpr int X_IMG = kNP * 128;         // one image (hi or lo) of the pixel tile: 256 rows x 128 B
constexpr int W_IMG = 128 * 128;         // one image of the weight k-block: 128 rows x 128 B
constexpr int STAGE = 2 * (X_IMG + W_IMG);
constexpr int NSTAGE = 2;
constexpr int SMEM = NSTAGE * STAGE + 1024 + 256;

struct P { const float* x; int x_ld; float* y; int y_ld; int M, nkb, ntiles; const float* wt; };

// weight image: [kb][hi,lo][128 rows = co][32 k] with the K-major SWIZZLE_128B chunk permutation (same as pack_tc_kernel)
__global__ void pack_w(const float* __restrict__ w /*[128][K]*/, int K, float* __restrict__ out) {
    const int nkb = K / 32;
    const long long total = (long long)nkb * 2 * 128 * 32;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        long long r = idx;
        const int e = (int)(r % 4); r /= 4;
        const int pc = (int)(r % 8); r /= 8;
        const int row = (int)(r % 128); r /= 128;
        const int op = (int)(r % 2); r /= 2;
        const int kb = (int)r;
        const int lc = pc ^ (row & 7);
        const float v = w[(size_t)row * K + kb * 32 + lc * 4 + e];
        const float hi = tf32_hi(v);
        out[idx] = op == 0 ? hi : tf32_hi(v - hi);
    }
}

__global__ void __launch_bounds__(kThreads, 1) conv1x1_t_kernel(const __grid_constant__ P p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
    const uint32_t bars = sbase + NSTAGE * STAGE;
    auto full_x = [&](int s) { return bars + 8u * s; };
    auto full_w = [&](int s) { return bars + 8u * (NSTAGE + s); };
    auto empty = [&](int s) { return bars + 8u * (2 * NSTAGE + s); };
    auto tmem_full = [&](int b) { return bars + 8u * (3 * NSTAGE + b); };
    auto tmem_empty = [&](int b) { return bars + 8u * (3 * NSTAGE + 2 + b); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sgen + NSTAGE * STAGE + 8 * (3 * NSTAGE + 4));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int NPW = kProd / 32, EPI0 = NPW, MMA_WARP = NPW + kEpi / 32, LOAD_WARP = MMA_WARP + 1;
    const int my_tiles = (p.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int nkb = p.nkb;
    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(full_x(s), NPW); mbar_init(full_w(s), 1); mbar_init(empty(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tmem_full(b), 1); mbar_init(tmem_empty(b), kEpi / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp < NPW) {
        // producers: 256 rows x 8 chunks = 2048 items per k-block, 4 per thread: rows rbase + 64*i
        const int chunk = tid & 7, rbase = tid >> 3;
        uint32_t s_off[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { const int r = rbase + 64 * i; s_off[i] = (uint32_t)r * 128u + (uint32_t)((chunk ^ (r & 7)) << 4); }
        const int total = my_tiles * nkb;
        int l_ti = 0, l_kb = 0;
        auto load_next = [&](float4 (&v)[4]) {
            if (l_ti >= my_tiles) return;
            const int tile = (int)blockIdx.x + l_ti * (int)gridDim.x;
            const float* base = p.x + (size_t)(tile * kNP + rbase) * p.x_ld + l_kb * 32 + chunk * 4;
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = __ldg(reinterpret_cast<const float4*>(base + (size_t)(64 * i) * p.x_ld));
            if (++l_kb == nkb) { l_kb = 0; ++l_ti; }
        };
        auto store_item = [&](int f, const float4 (&v)[4]) {
            const int s = f % NSTAGE; const uint32_t ph = (f / NSTAGE) & 1;
            mbar_wait(empty(s), ph ^ 1u);
            uint8_t* x_hi = sgen + s * STAGE;
            uint8_t* x_lo = x_hi + X_IMG;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 t = v[i];
                const float4 hi = make_float4(tf32_hi(t.x), tf32_hi(t.y), tf32_hi(t.z), tf32_hi(t.w));
                const float4 lo = make_float4(tf32_hi(t.x - hi.x), tf32_hi(t.y - hi.y), tf32_hi(t.z - hi.z), tf32_hi(t.w - hi.w));
                *reinterpret_cast<float4*>(x_hi + s_off[i]) = hi;
                *reinterpret_cast<float4*>(x_lo + s_off[i]) = lo;
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(full_x(s));
        };
        float4 va[4], vb[4];
        load_next(va);
        for (int f = 0; f < total; f += 2) {
            load_next(vb);
            store_item(f, va);
            if (f + 1 < total) { load_next(va); store_item(f + 1, vb); }
        }
    } else if (warp == MMA_WARP) {
        if (lane == 0) {
            // D=f32, A=B=tf32, both K-major, N=256, M=128
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kNP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            int f = 0;
            for (int ti = 0; ti < my_tiles; ++ti) {
                const int buf = ti & 1; const uint32_t tph = (ti >> 1) & 1;
                mbar_wait(tmem_empty(buf), tph ^ 1u);
                tc_fence_after();
                for (int kb = 0; kb < nkb; ++kb, ++f) {
                    const int s = f % NSTAGE; const uint32_t ph = (f / NSTAGE) & 1;
                    mbar_wait(full_x(s), ph);
                    mbar_wait(full_w(s), ph);
                    tc_fence_after();
                    const uint32_t x_hi = sbase + s * STAGE, x_lo = x_hi + X_IMG, w_hi = x_hi + 2 * X_IMG, w_lo = w_hi + W_IMG;
                    const uint32_t acc = tmem + (uint32_t)(buf * kNP);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const uint64_t dwh = make_desc(w_hi + kk * 32), dwl = make_desc(w_lo + kk * 32);
                        const uint64_t dxh = make_desc(x_hi + kk * 32), dxl = make_desc(x_lo + kk * 32);
                        mma_tf32(acc, dwl, dxh, idesc, (kb | kk) ? 1u : 0u);
                        mma_tf32(acc, dwh, dxl, idesc, 1u);
                        mma_tf32(acc, dwh, dxh, idesc, 1u);
                    }
                    mma_commit(empty(s));
                }
                mma_commit(tmem_full(buf));
            }
        }
        __syncwarp();
    } else if (warp == LOAD_WARP) {
        if (lane == 0) {
            constexpr uint32_t BYTES = 2 * W_IMG;
            int f = 0;
            for (int ti = 0; ti < my_tiles; ++ti)
                for (int kb = 0; kb < nkb; ++kb, ++f) {
                    const int s = f % NSTAGE; const uint32_t ph = (f / NSTAGE) & 1;
                    mbar_wait(empty(s), ph ^ 1u);
                    mbar_expect_tx(full_w(s), BYTES);
                    bulk_g2s(sbase + s * STAGE + 2 * X_IMG, reinterpret_cast<const uint8_t*>(p.wt) + (size_t)kb * BYTES, BYTES, full_w(s));
                }
        }
        __syncwarp();
    } else {
        // epilogue: lane quarter q = 32 output channels, `half` = 128 of the 256 pixel columns; for every pixel the warp's
        // 32 lanes write 32 consecutive floats
        const int q = warp & 3, half = (warp - EPI0) >> 2;
        const int co = q * 32 + lane;
        for (int ti = 0; ti < my_tiles; ++ti) {
            const int buf = ti & 1; const uint32_t tph = (ti >> 1) & 1;
            const int tile = (int)blockIdx.x + ti * (int)gridDim.x;
            mbar_wait(tmem_full(buf), tph);
            tc_fence_after();
            const uint32_t tb = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * kNP);
            float* ybase = p.y + (size_t)(tile * kNP) * p.y_ld + co;
            for (int c0 = half * 128; c0 < half * 128 + 128; c0 += 16) {
                float v[16];
                tmem_ld16(tb + (uint32_t)c0, v);
#pragma unroll
                for (int j = 0; j < 16; ++j) ybase[(size_t)(c0 + j) * p.y_ld] = v[j];
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty(buf));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

int main(int argc, char** argv) {
    const int M = argc > 1 ? atoi(argv[1]) : 262144, K = argc > 2 ? atoi(argv[2]) : 224, N = 128;
    if (M % kNP || K % 32) { printf("M %% 256 and K %% 32 must be 0\n"); return 1; }
    std::vector<float> hx((size_t)M * K), hw((size_t)N * K);
    srand(1);
    for (auto& v : hx) v = (float)rand() / RAND_MAX - 0.5f;
    for (auto& v : hw) v = ((float)rand() / RAND_MAX - 0.5f) * 0.2f;
    float *dx, *dw, *dwt, *dy;
    cudaMalloc(&dx, hx.size() * 4); cudaMalloc(&dw, hw.size() * 4); cudaMalloc(&dwt, (size_t)(K / 32) * 2 * 128 * 32 * 4); cudaMalloc(&dy, (size_t)M * N * 4);
    cudaMemcpy(dx, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dw, hw.data(), hw.size() * 4, cudaMemcpyHostToDevice);
    pack_w<<<256, 256>>>(dw, K, dwt);
    cudaFuncSetAttribute(conv1x1_t_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    P p{dx, K, dy, N, M, K / 32, M / kNP, dwt};
    const int grid = p.ntiles < 148 ? p.ntiles : 148;
    conv1x1_t_kernel<<<grid, kThreads, SMEM>>>(p);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<float> hy((size_t)512 * N);
    cudaMemcpy(hy.data(), dy, hy.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    for (int m = 0; m < 512; ++m) for (int n = 0; n < N; ++n) {
        double r = 0; for (int k = 0; k < K; ++k) r += (double)hx[(size_t)m * K + k] * hw[(size_t)n * K + k];
        maxerr = fmax(maxerr, fabs(r - hy[(size_t)m * N + n])); maxref = fmax(maxref, fabs(r));
    }
    printf("check (512 rows): max |err| / max |ref| = %.3e  (3xTF32 expects ~1e-6)\n", maxerr / maxref);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int i = 0; i < 20; ++i) conv1x1_t_kernel<<<grid, kThreads, SMEM>>>(p);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 20;
    printf("M %d K %d N %d: %.3f ms  %.1f TFLOP/s   (conv_tc.cu, same shape, no statistics: 0.123 ms / 122 TFLOP/s at M 262144 K 224)\n",
           M, K, N, ms, 2.0 * M * K * N / ms / 1e9);
    return 0;
}
