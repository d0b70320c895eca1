// conv_wgrad_pw.cu -- weight gradient of a pointwise (1x1, stride 1) convolution on tcgen05:
//
//   dw[cb][ca] = sum_pixels  Q[p][cb] * dY[p][ca]
//
// One CTA owns a contiguous pixel range and up to 512 Q channels x 128 dY channels of the result, which stay in TMEM
// (128 lanes = dY channels, <= 512 columns = Q channels) for the whole range: every activation is gathered, BN+ReLU
// transformed and tf32-split exactly ONCE (the generic kernel re-gathers dY for every 128-column tile of Q and spends
// ~85 instructions per 16-byte chunk in its producers).  Both operands are MN-major straight from NHWC memory:
// planes of 32 channels x 16 pixels, one 128-byte row per pixel (SWIZZLE_128B_BASE32B on absolute address bits).
// Per 8 pixels the tensor core gets M=128 x N<=256 x K=8 MMAs (3xTF32), the shape at which operand traffic from
// shared memory balances the tensor pipe.  DenseNet conv1 / transition / attention 1x1 layers.
//
//   warps 0-7  producers + epilogue (coalesced fp32 atomics: lane = dY channel);  warp 8  TMEM alloc + MMA issue.
#include "tc_common.cuh"

namespace saunet {

struct WgPwP {
    saunet_wgrad_desc d;
    long long M, pix_per_cta;
    int cb_per_cta;          // Q channels per CTA (multiple of 4, <= 32 * NP)
    int nstage;
};

constexpr int kPwProducers = 256;
constexpr int kPwThreads = kPwProducers + 32;
constexpr int kPwKP = 16;                                 // pixels per k-block (two K=8 MMA steps)
constexpr int kPwPlane = kPwKP * 128;                     // 2048: 32 channels x 16 pixels

template <int NP>                                         // 32-channel planes of Q per CTA (2, 4, 8, 16)
struct PwCfg {
    static constexpr int A_IMG = 4 * kPwPlane, B_IMG = NP * kPwPlane;
    static constexpr int STAGE = 2 * (A_IMG + B_IMG);
    static constexpr int A_ITEMS = kPwKP * 32 / kPwProducers;            // 2 float4 of dY per thread per k-block
    static constexpr int B_CH = NP * 8;                                  // 16-byte chunks per Q pixel
    static constexpr int B_PSTEP = kPwProducers / B_CH;                  // 16, 8, 4, 2
    static constexpr int B_ITEMS = kPwKP / B_PSTEP;                      // 1, 2, 4, 8
    static constexpr int MAX_STAGE = 4;
    static constexpr int TMEM_COLS = NP * 32 <= 32 ? 32 : (NP * 32 <= 64 ? 64 : (NP * 32 <= 128 ? 128 : (NP * 32 <= 256 ? 256 : 512)));
};

__device__ __forceinline__ uint32_t pw_swz(int cc, int px) {
    return (uint32_t)((((cc >> 1) ^ (px & 3)) << 5) | ((cc & 1) << 4));
}

template <int NP>
__global__ void __launch_bounds__(kPwThreads, 1) conv_wgrad_pw_kernel(const __grid_constant__ WgPwP p) {
    using Cfg = PwCfg<NP>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
    const int nstage = p.nstage;
    const uint32_t bars = sbase + nstage * Cfg::STAGE;
    auto full = [&](int s) { return bars + 8u * s; };
    auto empty = [&](int s) { return bars + 8u * (Cfg::MAX_STAGE + s); };
    const uint32_t accum_bar = bars + 8u * (2 * Cfg::MAX_STAGE);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sgen + nstage * Cfg::STAGE + 8 * (2 * Cfg::MAX_STAGE + 1));

    const saunet_wgrad_desc& d = p.d;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int MMA_WARP = kPwProducers / 32;
    const bool one = d.precision == 2;          // single-pass TF32: the lo images are neither written nor multiplied
    const int ca0 = blockIdx.y * 128;                     // dY channel tile (M side)
    const int cb0 = blockIdx.z * p.cb_per_cta;            // Q channel tile (N side)
    int ncb = d.Cb - cb0; if (ncb > p.cb_per_cta) ncb = p.cb_per_cta;
    const long long mbeg = (long long)blockIdx.x * p.pix_per_cta;
    long long mend = mbeg + p.pix_per_cta; if (mend > p.M) mend = p.M;
    const int nkb = (int)((mend - mbeg + kPwKP - 1) / kPwKP);          // host guarantees >= 1

    if (tid == 0) {
        for (int s = 0; s < nstage; ++s) { mbar_init(full(s), kPwProducers / 32); mbar_init(empty(s), 1); }
        mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(Cfg::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp < MMA_WARP) {
        // ================= producers =================
        // dY: chunk = tid & 31 (4 channels of the 128-wide tile), pixels (tid >> 5) + 8*i
        const int a_ch = tid & 31, a_px0 = tid >> 5;
        const int a_c = ca0 + a_ch * 4;
        const bool a_cv = a_c < d.Ca;
        const uint32_t a_s0 = (uint32_t)(a_ch >> 3) * kPwPlane + (uint32_t)a_px0 * 128u + pw_swz(a_ch & 7, a_px0);
        // Q: chunk = tid % B_CH, pixels tid / B_CH + B_PSTEP*i
        const int b_ch = tid % Cfg::B_CH, b_px0 = tid / Cfg::B_CH;
        const int b_c = cb0 + b_ch * 4;
        const bool b_cv = b_ch * 4 < ncb;
        const uint32_t b_pl = (uint32_t)(b_ch >> 3) * kPwPlane;
        float4 qsc = make_float4(1.f, 1.f, 1.f, 1.f), qsh = make_float4(0.f, 0.f, 0.f, 0.f);
        if (d.q_scale && b_cv) {
            qsc = __ldg(reinterpret_cast<const float4*>(d.q_scale + b_c));
            qsh = __ldg(reinterpret_cast<const float4*>(d.q_shift + b_c));
        }
        const float* pa = d.p + (size_t)(mbeg + a_px0) * d.p_ld + a_c;
        const float* pb = d.q + (size_t)(mbeg + b_px0) * d.q_ld + b_c;
        const size_t a_step = (size_t)8 * d.p_ld, b_step = (size_t)Cfg::B_PSTEP * d.q_ld;
        const size_t a_kb = (size_t)kPwKP * d.p_ld, b_kb = (size_t)kPwKP * d.q_ld;
        auto load_kb = [&](int kb, float4 (&va)[Cfg::A_ITEMS], float4 (&vb)[Cfg::B_ITEMS], unsigned& mask) {
            const long long m0 = mbeg + (long long)kb * kPwKP;
            unsigned mk = 0;
#pragma unroll
            for (int i = 0; i < Cfg::A_ITEMS; ++i) {
                va[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (a_cv && m0 + a_px0 + 8 * i < mend) va[i] = __ldg(reinterpret_cast<const float4*>(pa + kb * a_kb + i * a_step));
            }
#pragma unroll
            for (int i = 0; i < Cfg::B_ITEMS; ++i) {
                vb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (b_cv && m0 + b_px0 + Cfg::B_PSTEP * i < mend) { vb[i] = __ldg(reinterpret_cast<const float4*>(pb + kb * b_kb + i * b_step)); mk |= 1u << i; }
            }
            mask = mk;
        };
        auto split_store = [&](uint8_t* hi_img, uint8_t* lo_img, uint32_t off, const float4& v) {
            float4 hi = split_hi4(v);
            *reinterpret_cast<float4*>(hi_img + off) = hi;
            if (!one) *reinterpret_cast<float4*>(lo_img + off) = split_lo4(v, hi);
        };
        auto store_kb = [&](int kb, const float4 (&va)[Cfg::A_ITEMS], const float4 (&vb)[Cfg::B_ITEMS], unsigned mask) {
            const int s = kb % nstage; const uint32_t ph = (kb / nstage) & 1;
            mbar_wait(empty(s), ph ^ 1u);
            uint8_t* a_hi = sgen + s * Cfg::STAGE;
            uint8_t* a_lo = a_hi + Cfg::A_IMG;
            uint8_t* b_hi = a_lo + Cfg::A_IMG;
            uint8_t* b_lo = b_hi + Cfg::B_IMG;
            if (a_cv) {
#pragma unroll
                for (int i = 0; i < Cfg::A_ITEMS; ++i) split_store(a_hi, a_lo, a_s0 + (uint32_t)(i * 8 * 128), va[i]);
            }
            if (b_cv) {
#pragma unroll
                for (int i = 0; i < Cfg::B_ITEMS; ++i) {
                    float4 v = vb[i];
                    if (d.q_scale && (mask & (1u << i))) {          // pixels past the range stay exactly zero
                        v.x = fmaf(v.x, qsc.x, qsh.x); v.y = fmaf(v.y, qsc.y, qsh.y); v.z = fmaf(v.z, qsc.z, qsh.z); v.w = fmaf(v.w, qsc.w, qsh.w);
                        if (d.q_relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                    }
                    const int px = b_px0 + Cfg::B_PSTEP * i;          // (the swizzle phase px & 3 varies with i when B_PSTEP == 2)
                    split_store(b_hi, b_lo, b_pl + (uint32_t)px * 128u + pw_swz(b_ch & 7, px), v);
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(full(s));
        };
        float4 a0[Cfg::A_ITEMS], a1[Cfg::A_ITEMS], b0[Cfg::B_ITEMS], b1[Cfg::B_ITEMS];
        unsigned m0 = 0, m1 = 0;
        load_kb(0, a0, b0, m0);
        for (int kb = 0; kb < nkb; kb += 2) {
            if (kb + 1 < nkb) load_kb(kb + 1, a1, b1, m1);
            store_kb(kb, a0, b0, m0);
            if (kb + 1 < nkb) {
                if (kb + 2 < nkb) load_kb(kb + 2, a0, b0, m0);
                store_kb(kb + 1, a1, b1, m1);
            }
        }
        // ================= epilogue: TMEM (lane = dY channel, column = Q channel) -> dw[cb][ca] =================
        mbar_wait(accum_bar, 0u);
        tc_fence_after();
        const int q = warp & 3, half = warp >> 2;
        const int ca = ca0 + q * 32 + lane;
        float* dbase = d.dw + (size_t)cb0 * d.Ca + ca;
        for (int c0 = half * 16; c0 < ncb; c0 += 32) {
            float v[16];
            tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            if (ca < d.Ca) {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (c0 + j < ncb) atomicAdd(dbase + (size_t)(c0 + j) * d.Ca, v[j]);
            }
        }
        tc_fence_before();
    } else {
        // ================= MMA issuer: converged warp, one elected lane issues (see conv_halo_tma.cu) =================
        const bool leader = elect_one();
        {
            const uint64_t dT = ((uint64_t)(kPwPlane >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)1 << 61);
            const int ncol = (ncb + 15) & ~15;                       // N of the whole CTA, in MMAs of <= 256 columns
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % nstage; const uint32_t ph = (kb / nstage) & 1;
                mbar_wait(full(s), ph);
                tc_fence_after();
                const uint32_t a_hi = sbase + s * Cfg::STAGE;
                const uint32_t b_hi = a_hi + 2 * Cfg::A_IMG;
                if (leader) {
                for (int n0 = 0; n0 < ncol; n0 += 256) {
                    const int n = (ncol - n0) < 256 ? (ncol - n0) : 256;
                    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                                           ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
                    const uint32_t bq = b_hi + (uint32_t)(n0 >> 5) * kPwPlane;
#pragma unroll
                    for (int j = 0; j < 2; ++j) {                    // two 8-pixel K steps
                        const uint64_t dah = dT | (uint64_t)((a_hi + j * 1024) >> 4), dal = dT | (uint64_t)((a_hi + Cfg::A_IMG + j * 1024) >> 4);
                        const uint64_t dbh = dT | (uint64_t)((bq + j * 1024) >> 4), dbl = dT | (uint64_t)((bq + Cfg::B_IMG + j * 1024) >> 4);
                        if (!one) {
                            mma_tf32(tmem + (uint32_t)n0, dal, dbh, idesc, (kb | j) ? 1u : 0u);
                            mma_tf32(tmem + (uint32_t)n0, dah, dbl, idesc, 1u);
                        }
                        mma_tf32(tmem + (uint32_t)n0, dah, dbh, idesc, (one && !(kb | j)) ? 0u : 1u);
                    }
                }
                mma_commit(empty(s));
                }
                __syncwarp();
            }
            if (leader) mma_commit(accum_bar);
        }
        __syncwarp();
    }
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(Cfg::TMEM_COLS) : "memory");
    }
}

template <int NP>
static int launch_pw(WgPwP& p, dim3 grid, cudaStream_t st) {
    using Cfg = PwCfg<NP>;
    int nstage = (200 * 1024) / Cfg::STAGE;
    if (nstage > Cfg::MAX_STAGE) nstage = Cfg::MAX_STAGE;
    p.nstage = nstage;
    const int smem = nstage * Cfg::STAGE + 1024 + 256;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_wgrad_pw_kernel<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) { set_error("conv_wgrad_pw: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); return SAUNET_ERR_CUDA; }
        attr_set = true;
    }
    conv_wgrad_pw_kernel<NP><<<grid, kPwThreads, smem, st>>>(p);
    SAUNET_CHECK_LAUNCH("conv_wgrad_pw_kernel");
    return SAUNET_OK;
}

bool conv_wgrad_tc_eligible(const saunet_wgrad_desc* d);

bool conv_wgrad_pw_eligible(const saunet_wgrad_desc* d) {
    if (!conv_wgrad_tc_eligible(d)) return false;
    if (d->KH != 1 || d->KW != 1 || d->sy != 1 || d->sx != 1 || d->offy != 0 || d->offx != 0) return false;
    if (d->Hg != d->Hq || d->Wg != d->Wq) return false;
    // few-channel full-resolution layers have too little work per 16-pixel k-block (barrier round trips dominate):
    // they stay on the skinny / generic kernels
    if (d->Ca < 128 && d->Cb < 128) return false;
    return true;
}

int conv_wgrad_pw(const saunet_wgrad_desc* d, cudaStream_t st) {
    WgPwP p; p.d = *d;
    p.M = (long long)d->B * d->Hg * d->Wg;
    const int mtiles = cdiv(d->Ca, 128);
    const int ntiles = cdiv(d->Cb, 512);
    int cbt = cdiv(cdiv(d->Cb, ntiles), 4) * 4;                     // Q channels per CTA, balanced, multiple of 4
    p.cb_per_cta = cbt;
    // one CTA per SM; at least 64 pixels per CTA
    long long splits = kNumSMs / (mtiles * ntiles); if (splits < 1) splits = 1;
    long long maxs = (p.M + 63) / 64; if (splits > maxs) splits = maxs;
    long long ppc = (p.M + splits - 1) / splits;
    ppc = (ppc + kPwKP - 1) / kPwKP * kPwKP;
    splits = (p.M + ppc - 1) / ppc;
    p.pix_per_cta = ppc;
    dim3 grid((unsigned)splits, (unsigned)mtiles, (unsigned)ntiles);
    if (cbt <= 64) return launch_pw<2>(p, grid, st);
    if (cbt <= 128) return launch_pw<4>(p, grid, st);
    if (cbt <= 256) return launch_pw<8>(p, grid, st);
    return launch_pw<16>(p, grid, st);
}

}  // namespace saunet
