// Experiment: MN-major A operand whose four 32-channel atoms OVERLAP in shared memory (LBO = 128 bytes = one pixel):
// rows 32j..32j+31 of the M=128 operand are then the same 32 channels read j pixels further into a halo patch, i.e.
// three horizontally adjacent taps of a 3x3 weight gradient stacked along M.
// B[k][n] = (n == k) so D[32j + c][n] = patch[shift + j + n][c] for n < 8; patch[i][c] = i + c/64.
//   nvcc -gencode arch=compute_100a,code=sm_100a -I../../shape-attentive-unet_b200/csrc -o umma_mn_lbo_test umma_mn_lbo_test.cu
#include "tc_common.cuh"
#include <vector>
using namespace saunet;
namespace saunet { void set_error(const char*, ...) {} void count_launch(int) {} }

__device__ __forceinline__ uint64_t mkdesc_mn(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;
    return d;
}
__device__ __forceinline__ uint32_t pix_off(int px, int ch4) {     // flat pixel rows of 128 B, absolute-address swizzle
    const int pr = px & 3, cc = ch4 & 7;
    return (uint32_t)px * 128u + (uint32_t)((((cc >> 1) ^ pr) << 5) | ((cc & 1) << 4));
}

__global__ void __launch_bounds__(128) k(float* out, int shift, int lbo, int sbo) {
    extern __shared__ uint8_t raw[];
    const uint32_t sbase = (smem_u32(raw) + 1023u) & ~1023u;
    uint8_t* sg = raw + (sbase - smem_u32(raw));
    uint8_t* A = sg;                       // patch: 128 px x 32 ch = 16 KB
    uint8_t* B = sg + 16384;               // 8 px x 32 ch
    uint64_t* bar = (uint64_t*)(B + 4096);
    uint32_t* slot = (uint32_t*)(bar + 1);
    const int tid = threadIdx.x;
    for (int i = tid; i < 128 * 32; i += 128) {
        int px = i / 32, c = i % 32;
        *(float*)(A + pix_off(px, c / 4) + (c % 4) * 4) = (float)px + (float)c / 64.f;
    }
    for (int i = tid; i < 8 * 32; i += 128) {
        int px = i / 32, n = i % 32;
        *(float*)(B + pix_off(px, n / 4) + (n % 4) * 4) = (n == px) ? 1.f : 0.f;
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
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        mma_tf32(tmem, mkdesc_mn(sbase + shift * 128, lbo, sbo), mkdesc_mn(sbase + 16384, 512, 512), idesc, 0u);
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
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024);
    for (int lbo : {128, 256, 1280}) for (int sbo : {512, 1024}) for (int shift : {0, 1, 2, 10, 23}) {
        k<<<1, 128, 32 * 1024>>>(d, shift, lbo, sbo);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("lbo %d shift %d: CUDA error %s\n", lbo, shift, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
        int ok = 1;
        for (int m = 0; m < 128; ++m) for (int n = 0; n < 8; ++n) {
            int px = shift + (m / 32) * (lbo / 128) + (n / 4) * (sbo / 128) + (n % 4);
            if (px < 128 && fabsf(h[m * 32 + n] - ((float)px + (m % 32) / 64.f)) > 0.004f) ok = 0;
        }
        printf("lbo %4d sbo %4d shift %2d: %s | D[0][0..1] %.3f %.3f D[33][0..1] %.3f %.3f D[70][4] %.3f D[127][7] %.3f\n", lbo, sbo, shift, ok ? "OK " : "BAD",
               h[0], h[1], h[33 * 32], h[33 * 32 + 1], h[70 * 32 + 4], h[127 * 32 + 7]);
    }
    return 0;
}
