// Experiment: MN-major tf32 operands (SWIZZLE_128B_BASE32B) whose descriptor starts at a PIXEL offset that is not a
// multiple of the 4-pixel swizzle atom (a tap shift inside a halo patch of dY for the weight-gradient kernel).
// A[k][m] = (m == k) for the 8 pixels of one MMA, so D[m][n] = B[k = m][n] for m < 8.  The B patch holds
// B_full[i][n] = i + n/64 stored with the swizzle computed from the ABSOLUTE pixel index i.
//   nvcc -gencode arch=compute_100a,code=sm_100a -I../../shape-attentive-unet_b200/csrc -o umma_mn_shift_test umma_mn_shift_test.cu
#include "tc_common.cuh"
#include <vector>
using namespace saunet;
namespace saunet { void set_error(const char*, ...) {} void count_launch(int) {} }

__device__ __forceinline__ uint64_t mkdesc_mn(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t base_off) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(base_off & 7) << 49;
    d |= (uint64_t)1 << 61;
    return d;
}
__device__ __forceinline__ uint32_t mn_off(int px, int ch, int atoms_mn) {
    const int pr = px & 3, cc = ch & 7;
    return (uint32_t)(px >> 2) * (uint32_t)(atoms_mn * 512) + (uint32_t)(ch >> 3) * 512u + (uint32_t)pr * 128u +
           (uint32_t)((((cc >> 1) ^ pr) << 5) | ((cc & 1) << 4));
}

__global__ void __launch_bounds__(128) k(float* out, int shift, int sbo_b, int npix_mma) {
    extern __shared__ uint8_t raw[];
    const uint32_t sbase = (smem_u32(raw) + 1023u) & ~1023u;
    uint8_t* sg = raw + (sbase - smem_u32(raw));
    uint8_t* A = sg;                       // 8 px x 128 ch = 4096 B
    uint8_t* B = sg + 4096;                // patch: 64 px x 32 ch = 8192 B
    uint64_t* bar = (uint64_t*)(B + 8192);
    uint32_t* slot = (uint32_t*)(bar + 1);
    const int tid = threadIdx.x;
    for (int i = tid; i < 8 * 128; i += 128) {
        int px = i / 128, m = i % 128;
        *(float*)(A + mn_off(px, m / 4, 4) + (m % 4) * 4) = (m == px) ? 1.f : 0.f;
    }
    for (int i = tid; i < 64 * 32; i += 128) {
        int px = i / 32, n = i % 32;
        *(float*)(B + mn_off(px, n / 4, 1) + (n % 4) * 4) = (float)px + (float)n / 64.f;
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
        const uint32_t a0 = sbase, b0 = sbase + 4096 + shift * 128;
        mma_tf32(tmem, mkdesc_mn(a0, 512, 2048, 0), mkdesc_mn(b0, 512, sbo_b, 0), idesc, 0u);
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
    for (int sbo : {512, 1024, 1536}) for (int shift : {0, 1, 2, 3, 4, 5, 9, 10, 11}) {
        k<<<1, 128, 32 * 1024>>>(d, shift, sbo, 8);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("sbo %d shift %d: CUDA error %s\n", sbo, shift, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
        int ok = 1;
        for (int m = 0; m < 8; ++m) {
            int px = shift + (m / 4) * (sbo / 128) + (m % 4);
            for (int n = 0; n < 32; ++n) if (fabsf(h[m * 32 + n] - ((float)px + n / 64.f)) > 0.004f) ok = 0;
        }
        printf("sbo %4d shift %2d: %s | m0..7 ->", sbo, shift, ok ? "OK " : "BAD");
        for (int m = 0; m < 8; ++m) printf(" %.3f/%.3f", h[m * 32 + 1], h[m * 32 + 9]);
        printf("\n");
    }
    return 0;
}
