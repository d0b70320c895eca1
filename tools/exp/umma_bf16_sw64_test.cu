// Experiment: bf16 operands (tcgen05.mma.kind::f16) in K-major SWIZZLE_64B tiles -- 64-byte rows of 32 bf16 channels --
// with descriptors that start at an arbitrary ROW of a halo patch (64-byte granularity, i.e. not 128-byte aligned for odd
// shifts) and an arbitrary stride between 8-row groups.  A_full[r][0] = r & 255, A_full[r][k] = ((r & 7) * 32 + k) & 255 (all exact in bf16) stored swizzled by ABSOLUTE
// address (16-byte chunk index ^= address bits [7,9)); B = 32x32 identity, so D[m][n] = A[row(m)][n].
//   nvcc -gencode arch=compute_100a,code=sm_100a -I../../shape-attentive-unet_b200/csrc -o umma_bf16_sw64_test umma_bf16_sw64_test.cu
#include "tc_common.cuh"
#include <cuda_bf16.h>
#include <vector>
using namespace saunet;
namespace saunet { void set_error(const char*, ...) {} void count_launch(int) {} }

__device__ __forceinline__ uint64_t mkdesc64(uint32_t saddr, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;          // SWIZZLE_64B
    return d;
}
__device__ __forceinline__ void mma_bf16_(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_c), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}

__global__ void __launch_bounds__(128) k(float* out, int shift, int sbo) {
    extern __shared__ uint8_t raw[];
    const uint32_t sbase = (smem_u32(raw) + 1023u) & ~1023u;
    uint8_t* sg = raw + (sbase - smem_u32(raw));
    const int ROWS = 640;
    uint8_t* A = sg; uint8_t* B = sg + ROWS * 64;           // B at 40960 (1024-aligned)
    uint64_t* bar = (uint64_t*)(B + 32 * 64);
    uint32_t* slot = (uint32_t*)(bar + 1);
    const int tid = threadIdx.x;
    for (int i = tid; i < ROWS * 32; i += 128) {
        int r = i / 32, kk = i % 32, c = kk / 8, e = kk % 8;             // 16-byte chunk c (8 bf16), element e
        *(__nv_bfloat16*)(A + r * 64 + ((c ^ ((r >> 1) & 3)) << 4) + e * 2) = __float2bfloat16(kk == 0 ? (float)(r & 255) : (float)((((r & 7) << 5) + kk) & 255));
    }
    for (int i = tid; i < 32 * 32; i += 128) {
        int n = i / 32, kk = i % 32, c = kk / 8, e = kk % 8;
        *(__nv_bfloat16*)(B + n * 64 + ((c ^ ((n >> 1) & 3)) << 4) + e * 2) = __float2bfloat16((n == kk) ? 1.f : 0.f);
    }
    if (tid == 0) { mbar_init(smem_u32(bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(32) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = *slot;
    if (tid == 0) {
        // D=f32, A=B=bf16, K-major, N=32, M=128
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t a0 = sbase + shift * 64, b0 = sbase + ROWS * 64;
        for (int kk = 0; kk < 2; ++kk)
            mma_bf16_(tmem, mkdesc64(a0 + kk * 32, sbo), mkdesc64(b0 + kk * 32, 512), idesc, kk ? 1u : 0u);
        mma_commit(smem_u32(bar));
    }
    mbar_wait(smem_u32(bar), 0);
    tc_fence_after();
    const int warp = tid >> 5, lane = tid & 31;
    for (int c0 = 0; c0 < 32; c0 += 16) {
        float v[16];
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * 32 + c0 + j] = v[j];
    }
    tc_fence_before(); __syncthreads();
    if (tid < 32) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32) : "memory"); }
}

int main() {
    float* d; cudaMalloc(&d, 128 * 32 * 4);
    std::vector<float> h(128 * 32);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    int bad = 0;
    for (int sbo : {512, 640, 1024, 1280, 576}) for (int shift : {0, 1, 2, 3, 8, 11, 17, 20, 21, 22}) {
        k<<<1, 128, 64 * 1024>>>(d, shift, sbo);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("sbo %d shift %d: CUDA error %s\n", sbo, shift, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
        int ok = 1, okcols = 1;
        for (int m = 0; m < 128; ++m) {
            int expect_row = shift + (m / 8) * (sbo / 64) + (m % 8);
            if ((int)h[m * 32] != (expect_row & 255)) ok = 0;
            for (int n = 1; n < 32; ++n) if (h[m * 32 + n] != (float)((((expect_row & 7) << 5) + n) & 255)) okcols = 0;
        }
        if (!ok || !okcols) ++bad;
        printf("sbo %4d shift %2d: rows %s cols %s | m0..9 ->", sbo, shift, ok ? "OK " : "BAD", okcols ? "OK " : "BAD");
        for (int m = 0; m < 10; ++m) printf(" %.2f", h[m * 32 + 1]);
        printf("\n");
    }
    printf("%s\n", bad ? "SOME BAD" : "ALL OK");
    return 0;
}
