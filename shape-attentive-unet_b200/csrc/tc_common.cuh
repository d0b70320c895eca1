// tc_common.cuh -- mbarrier / tcgen05 / TMEM primitives shared by the tensor-core kernels (sm_100a inline PTX).
#pragma once
#include "common.cuh"

namespace saunet {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    long long t0 = 0;
    int spins = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return;
        // watchdog: a protocol bug must trap, not hang the device (clock read only every 4096 failed probes)
        if ((++spins & 4095) == 0) {
            if (t0 == 0) t0 = clock64();
            else if ((clock64() - t0) > 4000000000ll) __trap();
        }
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void mma_tf32(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_c), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// one lane of a CONVERGED warp (elect.sync): the lane that issues tcgen05 / bulk-copy work for the warp
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n\t"
        "tcgen05.wait::ld.sync.aligned;\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// round-to-nearest tf32 (low 13 mantissa bits zero afterwards): hi = rn(v), lo = rn(v - hi) keeps every dropped
// term of the 3xTF32 product at <= 2^-22 relative and unbiased
__device__ __forceinline__ float tf32_hi(float v) {
    // == cvt.rna.tf32.f32 for finite inputs, as two integer ops (the cvt pipe runs at quarter rate)
    return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u);
}


// Operand split of the activations (3xTF32): x = hi + lo with hi, lo both tf32.
// kind::tf32 reads only the top 19 bits of each 32-bit operand element (sign, 8 exponent, 10 mantissa bits): it
// TRUNCATES.  So the "hi" image can hold the fp32 value itself (hi := trunc(x), taken by the hardware for free) and
// lo := x - trunc(x) is exact in fp32 (<= 13 significant bits, again truncated to 11 by the hardware): 2 ALU ops per
// element instead of the 5 of a round-to-nearest split, and nothing to store for `hi` when the operand needs no
// prologue.  Errors: |x - hi - lo_used| <= 2^-21 |x| (one-sided), the dropped lo*lo term <= 2^-20 of the product.
// -DSAUNET_SPLIT_RN restores the round-to-nearest split (hi = rn(x), lo = rn(x - hi)).
#ifndef SAUNET_SPLIT_RN
__device__ __forceinline__ float split_lo1(float v) { return v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }
__device__ __forceinline__ float4 split_hi4(const float4& t) { return t; }
__device__ __forceinline__ float4 split_lo4(const float4& t, const float4&) {
    return make_float4(split_lo1(t.x), split_lo1(t.y), split_lo1(t.z), split_lo1(t.w));
}
#else
__device__ __forceinline__ float4 split_hi4(const float4& t) { return make_float4(tf32_hi(t.x), tf32_hi(t.y), tf32_hi(t.z), tf32_hi(t.w)); }
__device__ __forceinline__ float4 split_lo4(const float4& t, const float4& hi) {
    return make_float4(tf32_hi(t.x - hi.x), tf32_hi(t.y - hi.y), tf32_hi(t.z - hi.z), tf32_hi(t.w - hi.w));
}
#endif

// shared-memory matrix descriptor: K-major, SWIZZLE_128B, 128-byte rows, 8-row groups 1024 bytes apart
// same layout with an explicit stride between 8-row groups (halo patches: one image row of the patch per group)
__device__ __forceinline__ uint64_t make_desc_sbo(uint32_t saddr, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);          // start address
    d |= (uint64_t)1 << 16;                          // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset
    d |= (uint64_t)1 << 46;                          // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
    return d;
}
// ---- operand classes of the K-major kernels ----------------------------------------------------------------------------
// NPASS = 3: 3xTF32 (hi + lo images, 128-byte rows of 32 fp32 channels, SWIZZLE_128B), NPASS = 1: single-pass TF32 (hi image
// only), NPASS = kBF16: bf16 operands (kind::f16, fp32 accumulate) -- the same 32-channel chunk is a 64-byte row in the
// K-major SWIZZLE_64B pattern (16-byte chunk index ^= absolute address bits [7,9)), two K = 16 MMAs per chunk.  Like the
// 128-byte pattern it is a function of the absolute shared-memory address: a descriptor may start at any 64-byte row of a
// halo patch and the stride between 8-row groups may be any multiple of 64 (tools/exp/umma_bf16_sw64_test.cu, B200).
constexpr int kBF16 = 16;

template <int NPASS>
struct Opnd {
    static constexpr bool BF = NPASS == kBF16;
    static constexpr int ROW = BF ? 64 : 128;           // bytes per 32-channel operand row
    static constexpr int NOP = NPASS == 3 ? 2 : 1;      // images per operand
    static constexpr int KSTEPS = BF ? 2 : 4;           // MMAs (32 bytes of K each) per 32-channel chunk
    // byte offset, inside an image whose base is 1024-byte aligned, of the 4-channel piece `chunk` (0..7) of row r
    __device__ static __forceinline__ uint32_t off(int r, int chunk) {
        if (BF) return (uint32_t)r * 64u + (uint32_t)((((chunk >> 1) ^ ((r >> 1) & 3)) << 4) | ((chunk & 1) << 3));
        return (uint32_t)r * 128u + (uint32_t)((chunk ^ (r & 7)) << 4);
    }
    // store 4 channels of one row into the operand image(s)
    __device__ static __forceinline__ void store(uint8_t* hi_img, uint8_t* lo_img, uint32_t o, const float4& t) {
        if (BF) {
            uint32_t a, b;
            asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(a) : "f"(t.y), "f"(t.x));
            asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(b) : "f"(t.w), "f"(t.z));
            *reinterpret_cast<uint2*>(hi_img + o) = make_uint2(a, b);
        } else {
            const float4 hi = split_hi4(t);
            *reinterpret_cast<float4*>(hi_img + o) = hi;
            if (NPASS == 3) *reinterpret_cast<float4*>(lo_img + o) = split_lo4(t, hi);
        }
    }
    // shared-memory matrix descriptor (K-major), `sbo` = bytes between 8-row groups
    __device__ static __forceinline__ uint64_t desc(uint32_t saddr, uint32_t sbo = 8 * ROW) {
        uint64_t d = 0;
        d |= (uint64_t)((saddr >> 4) & 0x3FFF);
        d |= (uint64_t)1 << 16;
        d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
        d |= (uint64_t)1 << 46;
        d |= (uint64_t)(BF ? 4 : 2) << 61;               // SWIZZLE_64B : SWIZZLE_128B
        return d;
    }
    // instruction descriptor: D = f32, A = B = tf32 | bf16, both K-major, M = 128
    __device__ static __forceinline__ uint32_t idesc(int n) {
        const uint32_t fmt = BF ? 1u : 2u;
        return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    }
    __device__ static __forceinline__ void mma(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t id, uint32_t acc) {
        if (BF) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "setp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                ::"r"(tmem_c), "l"(adesc), "l"(bdesc), "r"(id), "r"(acc) : "memory");
        } else {
            mma_tf32(tmem_c, adesc, bdesc, id, acc);
        }
    }
};

// epilogue math on one 16-column chunk of a row: v = acc + bias (exactly 0 for rows / columns outside the problem, so
// that the BatchNorm statistics see nothing), o = act(v * rs).  Every branch is warp-uniform except the row mask.
__device__ __forceinline__ void epi_chunk(float (&v)[16], float (&o)[16], const float* bias, int nvalid, bool tile_full,
                                          bool mval, bool has_rs, float rs, int act) {
    if (bias) {
        if (nvalid >= 16 && ((reinterpret_cast<uintptr_t>(bias) & 15u) == 0)) {
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(bias) + q4);
                v[4 * q4] += b.x; v[4 * q4 + 1] += b.y; v[4 * q4 + 2] += b.z; v[4 * q4 + 3] += b.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) if (j < nvalid) v[j] += __ldg(bias + j);
        }
    }
    if (nvalid < 16) {
#pragma unroll
        for (int j = 0; j < 16; ++j) if (j >= nvalid) v[j] = 0.f;
    }
    if (!tile_full) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = mval ? v[j] : 0.f;
    }
    if (has_rs) {
#pragma unroll
        for (int j = 0; j < 16; ++j) o[j] = v[j] * rs;
    } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) o[j] = v[j];
    }
    if (act == SAUNET_ACT_RELU) {
#pragma unroll
        for (int j = 0; j < 16; ++j) o[j] = fmaxf(o[j], 0.f);
    } else if (act == SAUNET_ACT_SIGMOID) {
#pragma unroll
        for (int j = 0; j < 16; ++j) o[j] = sigmoid_acc(o[j]);
    }
}

// column sums of a [32 lanes][16 cols] register tile by transpose-reduce: 16 shuffles instead of 80.
// afterwards lane L holds the sum of column (L >> 1).
__device__ __forceinline__ float colsum16(float (&v)[16], int lane) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float send = (lane & 16) ? v[j] : v[j + 8], keep = (lane & 16) ? v[j + 8] : v[j];
        v[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float send = (lane & 8) ? v[j] : v[j + 4], keep = (lane & 8) ? v[j + 4] : v[j];
        v[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        float send = (lane & 4) ? v[j] : v[j + 2], keep = (lane & 4) ? v[j + 2] : v[j];
        v[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    {
        float send = (lane & 2) ? v[0] : v[1], keep = (lane & 2) ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}


}  // namespace saunet
