// conv_wgrad_halo.cu -- weight gradient of a 3x3 / stride 1 / pad 1 convolution on tcgen05, all 9 taps per CTA.
//
//   dw[tap][cb][ca] = sum_q  Q[q][cb] * dY[q - (tap - 1)][ca]          (dY = 0 outside the image)
//
// The generic kernel (conv_wgrad_tc.cu) gives every tap its own CTA tile, so Q is gathered, BN+ReLU-transformed
// and tf32-split 9 times and dY once per tile: its 16 producer warps are instruction-bound (ncu: 60 % issue-active,
// 12 % tensor-active on the DenseNet 3x3 layers).  Here one CTA walks 8x8-pixel tiles of the image and, per tile,
//   * stores the 64 Q pixels once (N side of the MMA, up to 128 channels, MN-major planes of 32 channels), and
//   * stores the 10x10 HALO patch of dY once (M side, planes of 32 channels, flat 128-byte pixel rows).
// The MN-major SWIZZLE_128B_BASE32B pattern is a function of the ABSOLUTE shared-memory address (bits [5,7) ^= bits
// [7,9)), so a descriptor may start at any pixel of the patch, and the four 32-row atoms of the M=128 operand may
// OVERLAP: with a leading-dimension byte offset of 128 (one pixel) rows 32j..32j+31 are the same 32 dY channels read
// j pixels further right, i.e. the three horizontally adjacent taps kx = 2-j stacked along M (the 4th slot is
// ignored).  (tools/exp/umma_mn_shift_test.cu, umma_mn_lbo_test.cu: verified on B200.)  One MMA per (ky, tile row)
// therefore covers 3 taps x 32 dY channels x up to 128 Q channels x 8 pixels, which balances the tensor pipe
// against the shared-memory operand bandwidth (the per-tap N=32 formulation is bound by re-reading the Q tile).
// 3 (ky) x planes accumulators of [128 x NT] fp32 live in TMEM for the CTA's whole pixel range; coalesced fp32
// atomics at the end.
//
//   warps 0-7  producers (128-bit gathers, fused BN+ReLU prologue recompute on Q, hi/lo split, swizzled stores),
//              then the epilogue;   warp 8  TMEM alloc + single-lane MMA issue.
#include "tc_common.cuh"

namespace saunet {

struct WgHaloP {
    saunet_wgrad_desc d;
    int tiles_x, tiles_y, ntiles, tiles_per_cta;
    int nstage;
};

constexpr int kWhProducers = 256;
constexpr int kWhThreads = kWhProducers + 32;
constexpr int kWhPatch = 100;                                  // 10 x 10 halo pixels
constexpr int kWhBPlane = (kWhPatch * 128 + 511) / 512 * 512;  // 13312: one 32-channel plane of the patch
constexpr int kWhAPlane = 64 * 128;                            // 8192: one 32-channel plane of the 64-pixel tile

template <int NB, int MA>
struct WhCfg {
    static constexpr int NBP = NB / 32;                        // dY planes
    static constexpr int A_IMG = MA * kWhAPlane, B_IMG = NBP * kWhBPlane;
    static constexpr int STAGE = 2 * (A_IMG + B_IMG);          // hi + lo
    static constexpr int NT = MA * 32;                         // Q channels per CTA = MMA N
    static constexpr int NACC = 3 * NBP;                       // accumulators: (dY plane, ky), [128 x NT] each
    static_assert(NACC * NT <= 512, "TMEM columns");
    static constexpr int A_ITEMS = 2 * MA;                     // float4 per producer thread per tile (Q)
    static constexpr int A_CH = MA * 8;                        // 16-byte chunks per Q pixel
    static constexpr int A_PSTEP = kWhProducers / A_CH;        // pixel step between a thread's items (8, 16, 32)
    static constexpr int B_CH = NB / 4;                        // chunks per dY pixel
    static constexpr int B_PSTEP = kWhProducers / B_CH;        // 32 or 16
    static constexpr int B_ITEMS = (kWhPatch + B_PSTEP - 1) / B_PSTEP;   // 4 or 7
    static constexpr int MAX_STAGE = 4;
};

__device__ __forceinline__ uint32_t wh_swz(int cc, int px) {          // 16-byte chunk cc (0..7) of pixel px inside its 128-byte row
    return (uint32_t)((((cc >> 1) ^ (px & 3)) << 5) | ((cc & 1) << 4));
}

template <int NB, int MA>
__global__ void __launch_bounds__(kWhThreads, 1) conv_wgrad_halo_kernel(const __grid_constant__ WgHaloP p) {
    using Cfg = WhCfg<NB, MA>;
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
    constexpr int MMA_WARP = kWhProducers / 32;
    const bool one = d.precision == 2;          // single-pass TF32: the lo images are neither written nor multiplied
    const int cb0 = blockIdx.y * Cfg::NT;                // first Q channel of this CTA's tile
    const int t_beg = blockIdx.x * p.tiles_per_cta;
    int t_end = t_beg + p.tiles_per_cta; if (t_end > p.ntiles) t_end = p.ntiles;
    const int nt = t_end - t_beg;                        // host guarantees nt >= 1

    if (tid == 0) {
        for (int s = 0; s < nstage; ++s) { mbar_init(full(s), kWhProducers / 32); mbar_init(empty(s), 1); }
        mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp < MMA_WARP) {
        // ================= producers =================
        // Q items: chunk (fixed per thread) of pixels a_px0 + A_PSTEP*i; the pixel's column inside the tile is fixed,
        // its row advances by A_PSTEP/8 per item.
        const int a_ch = tid % Cfg::A_CH, a_px0 = tid / Cfg::A_CH;
        const int a_c = cb0 + a_ch * 4;
        const bool a_cv = a_c < d.Cb;
        const int a_g0 = ((a_px0 >> 3) * d.Wq + (a_px0 & 7)) * d.q_ld + a_c;       // element offset relative to the tile origin
        const int a_gstep = (Cfg::A_PSTEP >> 3) * d.Wq * d.q_ld;
        const uint32_t a_s0 = (uint32_t)(a_ch >> 3) * kWhAPlane + (uint32_t)a_px0 * 128u + wh_swz(a_ch & 7, a_px0);
        float4 qsc = make_float4(1.f, 1.f, 1.f, 1.f), qsh = make_float4(0.f, 0.f, 0.f, 0.f);
        if (d.q_scale && a_cv) {
            qsc = __ldg(reinterpret_cast<const float4*>(d.q_scale + a_c));
            qsh = __ldg(reinterpret_cast<const float4*>(d.q_shift + a_c));
        }
        // dY items: chunk (fixed) of patch pixels b_px0 + B_PSTEP*i
        const int b_ch = tid % Cfg::B_CH, b_px0 = tid / Cfg::B_CH;
        const int b_c = b_ch * 4;
        const bool b_cv = b_c < d.Ca;
        const uint32_t b_s0 = (uint32_t)(b_ch >> 3) * kWhBPlane + (uint32_t)b_px0 * 128u + wh_swz(b_ch & 7, b_px0);
        int b_rc[Cfg::B_ITEMS];            // patch (row << 8 | col), -1: no such item
#pragma unroll
        for (int i = 0; i < Cfg::B_ITEMS; ++i) {
            const int pp = b_px0 + Cfg::B_PSTEP * i;
            b_rc[i] = pp < kWhPatch ? (((pp / 10) << 8) | (pp % 10)) : -1;
        }
        auto tile_origin = [&](int t, int& b, int& y0, int& x0) {
            const int txi = t % p.tiles_x; t /= p.tiles_x;
            const int tyi = t % p.tiles_y; b = t / p.tiles_y;
            y0 = tyi * 8; x0 = txi * 8;
        };
        auto load_tile = [&](int t, float4 (&va)[Cfg::A_ITEMS], float4 (&vb)[Cfg::B_ITEMS]) {
            int b, y0, x0; tile_origin(t, b, y0, x0);
            const float* qb = d.q + ((b * d.Hq + y0) * d.Wq + x0) * d.q_ld + a_g0;
#pragma unroll
            for (int i = 0; i < Cfg::A_ITEMS; ++i) {
                va[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (a_cv) va[i] = __ldg(reinterpret_cast<const float4*>(qb + i * a_gstep));
            }
#pragma unroll
            for (int i = 0; i < Cfg::B_ITEMS; ++i) {
                vb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                const int iy = y0 - 1 + (b_rc[i] >> 8), ix = x0 - 1 + (b_rc[i] & 255);
                if (b_rc[i] >= 0 && b_cv && (unsigned)iy < (unsigned)d.Hq && (unsigned)ix < (unsigned)d.Wq)
                    vb[i] = __ldg(reinterpret_cast<const float4*>(d.p + ((b * d.Hq + iy) * d.Wq + ix) * d.p_ld + b_c));
            }
        };
        auto split_store = [&](uint8_t* hi_img, uint8_t* lo_img, uint32_t off, const float4& v) {
            float4 hi = split_hi4(v);
            *reinterpret_cast<float4*>(hi_img + off) = hi;
            if (!one) *reinterpret_cast<float4*>(lo_img + off) = split_lo4(v, hi);
        };
        auto store_tile = [&](int it, const float4 (&va)[Cfg::A_ITEMS], const float4 (&vb)[Cfg::B_ITEMS]) {
            const int s = it % nstage; const uint32_t ph = (it / nstage) & 1;
            mbar_wait(empty(s), ph ^ 1u);
            uint8_t* a_hi = sgen + s * Cfg::STAGE;
            uint8_t* a_lo = a_hi + Cfg::A_IMG;
            uint8_t* b_hi = a_lo + Cfg::A_IMG;
            uint8_t* b_lo = b_hi + Cfg::B_IMG;
            if (a_cv) {
#pragma unroll
                for (int i = 0; i < Cfg::A_ITEMS; ++i) {
                    float4 v = va[i];
                    if (d.q_scale) {
                        v.x = fmaf(v.x, qsc.x, qsh.x); v.y = fmaf(v.y, qsc.y, qsh.y); v.z = fmaf(v.z, qsc.z, qsh.z); v.w = fmaf(v.w, qsc.w, qsh.w);
                        if (d.q_relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                    }
                    split_store(a_hi, a_lo, a_s0 + (uint32_t)(i * Cfg::A_PSTEP * 128), v);
                }
            }
            if (b_cv) {
#pragma unroll
                for (int i = 0; i < Cfg::B_ITEMS; ++i)
                    if (b_rc[i] >= 0) split_store(b_hi, b_lo, b_s0 + (uint32_t)(i * Cfg::B_PSTEP * 128), vb[i]);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(full(s));
        };
        float4 a0[Cfg::A_ITEMS], a1[Cfg::A_ITEMS], b0[Cfg::B_ITEMS], b1[Cfg::B_ITEMS];
        load_tile(t_beg, a0, b0);
        for (int it = 0; it < nt; it += 2) {
            if (it + 1 < nt) load_tile(t_beg + it + 1, a1, b1);
            store_tile(it, a0, b0);
            if (it + 1 < nt) {
                if (it + 2 < nt) load_tile(t_beg + it + 2, a0, b0);
                store_tile(it + 1, a1, b1);
            }
        }
        // ================= epilogue: TMEM -> coalesced fp32 atomics into dw[(tap, cb)][ca] =================
        // accumulator (plane, ky): lane quarter j holds tap kx = 2 - j, lane = dY channel, column = Q channel
        mbar_wait(accum_bar, 0u);
        tc_fence_after();
        const int q = warp & 3, half = warp >> 2;
        if (q < 3) {
            const int kx = 2 - q;
            int item = 0;
            for (int acc = 0; acc < Cfg::NACC; ++acc) {
                const int pl = acc / 3, ky = acc - pl * 3;
                const int ca = pl * 32 + lane;
                float* dbase = d.dw + ((size_t)(ky * 3 + kx) * d.Cb + cb0) * d.Ca + ca;
#pragma unroll 1
                for (int c0 = 0; c0 < Cfg::NT; c0 += 16, ++item) {
                    if ((item & 1) != half || cb0 + c0 >= d.Cb) continue;      // warp-uniform
                    float v[16];
                    tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * Cfg::NT + c0), v);
                    if (ca < d.Ca) {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (cb0 + c0 + j < d.Cb) atomicAdd(dbase + (size_t)(c0 + j) * d.Ca, v[j]);
                    }
                }
            }
        }
        tc_fence_before();
    } else {
        // ================= MMA issuer: converged warp, one elected lane issues (see conv_halo_tma.cu) =================
        const bool leader = elect_one();
        {
            // D=f32, A=B=tf32, both MN-major, N=NT, M=128
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                                   ((uint32_t)(Cfg::NT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            // descriptor templates.  M side (dY patch): LBO = 128 -> the four 32-row atoms are the patch read 0..3 pixels
            // further right; N side (Q tile): LBO = distance between 32-channel planes.  SBO = 512: next 4 pixels of the row.
            const uint64_t dP = ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)1 << 61);
            const uint64_t dQ = ((uint64_t)(kWhAPlane >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)1 << 61);
            for (int it = 0; it < nt; ++it) {
                const int s = it % nstage; const uint32_t ph = (it / nstage) & 1;
                mbar_wait(full(s), ph);
                tc_fence_after();
                const uint32_t q_hi = sbase + s * Cfg::STAGE;
                const uint64_t dqh0 = dQ | (uint64_t)(q_hi >> 4), dql0 = dQ | (uint64_t)((q_hi + Cfg::A_IMG) >> 4);
                const uint32_t p_hi = q_hi + 2 * Cfg::A_IMG;
                if (leader) {
#pragma unroll 1
                for (int acc = 0; acc < Cfg::NACC; ++acc) {
                    const int pl = acc / 3, ky = acc - pl * 3;
                    const uint32_t pimg = p_hi + pl * kWhBPlane + (uint32_t)((2 - ky) * 10 * 128);
                    const uint64_t dph0 = dP | (uint64_t)(pimg >> 4), dpl0 = dP | (uint64_t)((pimg + Cfg::B_IMG) >> 4);
                    const uint32_t tacc = tmem + (uint32_t)(acc * Cfg::NT);
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        const uint64_t po = (uint64_t)(r * 1280 >> 4), qo = (uint64_t)(r * 1024 >> 4);
                        if (!one) {
                            mma_tf32(tacc, dpl0 + po, dqh0 + qo, idesc, (it | r) ? 1u : 0u);
                            mma_tf32(tacc, dph0 + po, dql0 + qo, idesc, 1u);
                        }
                        mma_tf32(tacc, dph0 + po, dqh0 + qo, idesc, (one && !(it | r)) ? 0u : 1u);
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
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

template <int NB, int MA>
static int launch_wh(WgHaloP& p, cudaStream_t st) {
    using Cfg = WhCfg<NB, MA>;
    const saunet_wgrad_desc& d = p.d;
    int nstage = (200 * 1024) / Cfg::STAGE;
    if (nstage > Cfg::MAX_STAGE) nstage = Cfg::MAX_STAGE;
    if (nstage < 2) { set_error("conv_wgrad_halo: stage too large"); return SAUNET_ERR_BAD_SHAPE; }
    p.nstage = nstage;
    const int smem = nstage * Cfg::STAGE + 1024 + 256;
    static int attr_smem = 0;
    if (attr_smem < smem) {
        cudaError_t e = cudaFuncSetAttribute(conv_wgrad_halo_kernel<NB, MA>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) { set_error("conv_wgrad_halo: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); return SAUNET_ERR_CUDA; }
        attr_smem = smem;
    }
    const int mtiles = cdiv(d.Cb, Cfg::NT);
    // one CTA per SM; cap the pixel range of a CTA (the tensor core truncates its fp32 accumulator on every MMA, a bias
    // that grows with the length of the accumulation chain)
    int ctas = kNumSMs / mtiles; if (ctas < 1) ctas = 1;
    const int waves = cdiv(p.ntiles, (long long)ctas * 48);            // <= 48 tiles (3072 pixels) per chain
    int tpc = cdiv(p.ntiles, (long long)ctas * waves); if (tpc < 1) tpc = 1;
    p.tiles_per_cta = tpc;
    dim3 grid(cdiv(p.ntiles, tpc), mtiles, 1);
    conv_wgrad_halo_kernel<NB, MA><<<grid, kWhThreads, smem, st>>>(p);
    SAUNET_CHECK_LAUNCH("conv_wgrad_halo_kernel");
    return SAUNET_OK;
}

bool conv_wgrad_tc_eligible(const saunet_wgrad_desc* d);

bool conv_wgrad_halo_eligible(const saunet_wgrad_desc* d) {
    if (!conv_wgrad_tc_eligible(d)) return false;
    if (d->KH != 3 || d->KW != 3 || d->sy != 1 || d->sx != 1 || d->offy != -1 || d->offx != -1) return false;
    if (d->Hg != d->Hq || d->Wg != d->Wq || d->Hq % 8 || d->Wq % 8) return false;
    if (d->Ca > 64) return false;
    if (!aligned16(d->dw)) return false;
    return true;
}

int conv_wgrad_halo(const saunet_wgrad_desc* d, cudaStream_t st) {
    WgHaloP p; p.d = *d;
    p.tiles_x = d->Wq / 8; p.tiles_y = d->Hq / 8; p.ntiles = d->B * p.tiles_x * p.tiles_y;
    if (d->Ca <= 32) {                                 // one dY plane: Q tiles of up to 128 channels
        if (d->Cb > 64) return launch_wh<32, 4>(p, st);
        if (d->Cb > 32) return launch_wh<32, 2>(p, st);
        return launch_wh<32, 1>(p, st);
    }
    if (d->Cb > 32) return launch_wh<64, 2>(p, st);    // two dY planes: 6 accumulators -> Q tiles of 64 channels
    return launch_wh<64, 1>(p, st);
}

}  // namespace saunet
