// Experiment: can a K-major SWIZZLE_128B UMMA descriptor start at a row offset that is NOT 1024-byte aligned
// (im2col tap shift inside a shared-memory halo patch)?  A_full[r][k] = r + k/64 is stored swizzled by ABSOLUTE row
// index; B = 32x32 identity, so D[m][n] = A[row(m)][n].  Prints which source row each output row m came from.
//   nvcc -gencode arch=compute_100a,code=sm_100a -I../../shape-attentive-unet_b200/csrc -o umma_shift_test umma_shift_test.cu
#include "tc_common.cuh"
#include <vector>
using namespace saunet;
namespace saunet { void set_error(const char*, ...) {} void count_launch(int) {} }

__device__ __forceinline__ uint64_t mkdesc(uint32_t saddr, uint32_t sbo, uint32_t base_off) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(base_off & 7) << 49;
    d |= (uint64_t)2 << 61;
    return d;
}

__global__ void __launch_bounds__(128) k(float* out, int shift, int sbo, int use_base_off) {
    extern __shared__ uint8_t raw[];
    const uint32_t sbase = (smem_u32(raw) + 1023u) & ~1023u;
    uint8_t* sg = raw + (sbase - smem_u32(raw));
    const int ROWS = 320;                        // enough rows for sbo = 2048 (16-row pitch)
    uint8_t* A = sg; uint8_t* B = sg + ROWS * 128;          // B at 40960 (1024-aligned)
    uint64_t* bar = (uint64_t*)(B + 32 * 128);
    uint32_t* slot = (uint32_t*)(bar + 1);
    const int tid = threadIdx.x;
    for (int i = tid; i < ROWS * 32; i += 128) {
        int r = i / 32, kk = i % 32, c = kk / 4, e = kk % 4;
        *(float*)(A + r * 128 + ((c ^ (r & 7)) << 4) + e * 4) = (float)(r & 15) + (float)kk / 64.f;
    }
    for (int i = tid; i < 32 * 32; i += 128) {
        int n = i / 32, kk = i % 32, c = kk / 4, e = kk % 4;
        *(float*)(B + n * 128 + ((c ^ (n & 7)) << 4) + e * 4) = (n == kk) ? 1.f : 0.f;
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
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t a0 = sbase + shift * 128, b0 = sbase + ROWS * 128;
        for (int kk = 0; kk < 4; ++kk)
            mma_tf32(tmem, mkdesc(a0 + kk * 32, sbo, use_base_off ? (a0 >> 7) : 0), mkdesc(b0 + kk * 32, 1024, 0), idesc, kk ? 1u : 0u);
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
    for (int sbo : {1024, 2048, 1280, 1152}) for (int bo = 0; bo < 1; ++bo) for (int shift : {0, 1, 3, 8, 17}) {
        k<<<1, 128, 64 * 1024>>>(d, shift, sbo, bo);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("sbo %d base_off %d shift %d: CUDA error %s\n", sbo, bo, shift, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
        int ok = 1, okcols = 1;
        for (int m = 0; m < 128; ++m) {
            int expect_row = shift + (m / 8) * (sbo / 128) + (m % 8);
            if ((int)h[m * 32] != (expect_row & 15)) ok = 0;
            for (int n = 0; n < 32; ++n) if (fabsf(h[m * 32 + n] - ((float)(expect_row & 15) + n / 64.f)) > 0.004f) okcols = 0;
        }
        printf("sbo %4d base_off %d shift %2d: rows %s cols %s | m0..9 ->", sbo, bo, shift, ok ? "OK " : "BAD", okcols ? "OK " : "BAD");
        for (int m = 0; m < 10; ++m) printf(" %.2f", h[m * 32 + 1]);
        printf("\n");
    }
    return 0;
}
