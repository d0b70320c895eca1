// conv_wgrad_tc.cu -- convolution weight gradient on tcgen05 tensor cores.
//
//   dw[(tap, cb)][ca] += sum_pixels  Q[gather(pixel, tap)][cb] * P[pixel][ca]
//
// is a GEMM whose reduction dimension is the PIXEL axis, so both operands are "MN-major" for the tensor core:
// NHWC memory already has the channels of one pixel contiguous, and a 128-byte row of 32 channels of one pixel is
// exactly one row of the UMMA MN-major SWIZZLE_128B atom (32 channels x 8 pixels).  No transposes anywhere.
//
// One CTA owns (tap, 128-channel tile of the M-side operand, BN-channel tile of the N-side operand) and a
// contiguous range of pixels (split-K over the grid x axis); fp32 partial tiles are combined with atomics.
// The host picks which of Q / P sits on the 128-row M side (`swap`), so that the wider operand fills M.
//   warps 0-7  producers: 128-bit gathers (im2col shift + padding for Q, fused BN+ReLU prologue recompute),
//              tf32 hi/lo split, MN-major swizzled stores, fence.proxy.async, arrive.   Then the epilogue.
//   warp 8     TMEM alloc + single-lane tcgen05.mma.kind::tf32 issue (M=128, N=BN, K=8 pixels), 3xTF32.
#include "tc_common.cuh"

namespace saunet {

struct WgTcP {
    saunet_wgrad_desc d;
    long long M;            // pixels of the output grid
    int HgWg;
    long long pix_per_split;
    int swap;               // 0: M side = Q (cb), N side = P (ca);  1: M side = P (ca), N side = Q (cb)
    int mtiles, ntiles;
};

// MN-major tf32 operands have exactly one legal shared-memory layout: SWIZZLE_128B_BASE32B (layout type 1).
// Atom = 4 pixels (K) x 128 bytes (32 channels, MN); inside a row the four 32-byte granules are XOR-swizzled with
// the row index (byte-address bits [5,7) ^= bits [7,9)).  LBO = distance between 32-channel atoms, SBO = distance
// between 4-pixel atoms; one K=8 MMA spans two of them.
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;
    return d;
}
// byte offset of 16-byte chunk `ch` (4 channels) of pixel `px` inside a tile of `atoms_mn` 32-channel atoms
__device__ __forceinline__ uint32_t mn_off(int px, int ch, int atoms_mn) {
    const int pr = px & 3, cc = ch & 7;
    return (uint32_t)(px >> 2) * (uint32_t)(atoms_mn * 512) + (uint32_t)(ch >> 3) * 512u + (uint32_t)pr * 128u +
           (uint32_t)((((cc >> 1) ^ pr) << 5) | ((cc & 1) << 4));
}

template <int BN>
struct WgCfg {
    static constexpr int KB = 32;                                  // pixels per k-block
    static constexpr int R_BYTES = KB * 128 * 4;                   // one image of the M-side tile (16 KB)
    static constexpr int S_BYTES = KB * BN * 4;
    static constexpr int STAGE = 2 * (R_BYTES + S_BYTES);          // hi + lo of both
    static constexpr int NSTAGE_RAW = (192 * 1024) / STAGE;
    static constexpr int MIN_CTAS = 1;       // (two CTAs/SM would cap registers at 113 and spill the 3-deep gather ring)
    static constexpr int NSTAGE = MIN_CTAS == 2 ? 2 : (NSTAGE_RAW > 4 ? 4 : NSTAGE_RAW);
    static constexpr int SMEM = NSTAGE * STAGE + 1024 + 256;
    static constexpr int NACC = ((512 / MIN_CTAS) / BN) > 4 ? 4 : ((512 / MIN_CTAS) / BN);
    static constexpr int TMEM_COLS = NACC * BN;
};

constexpr int kWgProducers = 512;      // 16 producer warps: the transform is issue-latency bound, 4 warps/scheduler hide it
constexpr int kWgThreads = kWgProducers + 32;

// source operand description seen by the producers
struct OpSrc {
    const float* ptr; int ld; int C; int c0;      // channel tile start
    int gather;                                   // 1: Q (shifted by the tap, zero outside the image), 0: P
};

template <int BN>
__global__ void __launch_bounds__(kWgThreads, WgCfg<BN>::MIN_CTAS) conv_wgrad_tc_kernel(const __grid_constant__ WgTcP p) {
    using Cfg = WgCfg<BN>;
    constexpr int NSTAGE = Cfg::NSTAGE, NACC = Cfg::NACC;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
    const uint32_t bars = sbase + NSTAGE * Cfg::STAGE;
    auto full = [&](int s) { return bars + 8u * s; };
    auto empty = [&](int s) { return bars + 8u * (NSTAGE + s); };
    const uint32_t accum_bar = bars + 8u * (2 * NSTAGE);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sgen + NSTAGE * Cfg::STAGE + 8 * (2 * NSTAGE + 1));

    const saunet_wgrad_desc& d = p.d;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int MMA_WARP = kWgProducers / 32;
    const bool one = d.precision == 2;          // single-pass TF32: the lo images are neither written nor multiplied

    // tile decode.  The gathered operand Q is addressed by kq = tap*Cb + cb (im2col column), so a 128-wide tile
    // spans several taps when Cb is small and dY (P) is read once per tile, not once per tap.
    const int nt = blockIdx.y % p.ntiles, mt = blockIdx.y / p.ntiles;
    const int CQ = d.KH * d.KW * d.Cb;
    OpSrc R, S;
    if (!p.swap) { R = {d.q, d.q_ld, CQ, mt * 128, 1}; S = {d.p, d.p_ld, d.Ca, nt * BN, 0}; }
    else         { R = {d.p, d.p_ld, d.Ca, mt * 128, 0}; S = {d.q, d.q_ld, CQ, nt * BN, 1}; }

    const long long mbeg = (long long)blockIdx.x * p.pix_per_split;
    long long mend = mbeg + p.pix_per_split; if (mend > p.M) mend = p.M;
    const int nkb = mend > mbeg ? (int)((mend - mbeg + Cfg::KB - 1) / Cfg::KB) : 0;

    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(full(s), kWgProducers / 32); mbar_init(empty(s), 1); }
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
        if (nkb > 0) {
            // ---------------- producers ----------------
            // M side: 32 chunks (128 columns) per pixel: chunk = tid & 31, pixels (tid >> 5) + RSTEP*i, i < RIT
            // N side: BN/4 chunks per pixel; when the tile has fewer than 512 chunks only the first threads load it
            constexpr int RSTEP = kWgProducers / 32, RIT = Cfg::KB / RSTEP;       // 16, 2
            constexpr int SCH = BN / 4;                       // chunks per pixel on the N side
            constexpr int SPIX_RAW = kWgProducers / SCH;
            constexpr int SPIX = SPIX_RAW > Cfg::KB ? Cfg::KB : SPIX_RAW;          // pixels covered per pass
            constexpr int SIT = Cfg::KB / SPIX;               // passes (>= 1)
            const bool s_active = tid < SCH * SPIX;
            const int rch = tid & 31, rp0 = tid >> 5;
            const int sch = tid % SCH, sp0 = tid / SCH;
            const int rk = R.c0 + rch * 4, sk = S.c0 + sch * 4;          // logical column of this thread's chunk
            const bool rcv = rk < R.C, scv = s_active && sk < S.C;
            // the gathered side's column -> (tap, channel); fixed per thread
            const int gk = p.swap ? sk : rk;
            const bool gv = p.swap ? scv : rcv;
            int gtap = 0, gc = 0;
            if (gv) { gtap = gk / d.Cb; gc = gk - gtap * d.Cb; }
            const int gky = gtap / d.KW, gkx = gtap - gky * d.KW;
            const int rc = R.gather ? gc : rk, sc_ = S.gather ? gc : sk;   // channel offset inside the source pixel
            float4 qsc = make_float4(1.f, 1.f, 1.f, 1.f), qsh = make_float4(0.f, 0.f, 0.f, 0.f);
            if (d.q_scale && gv) {
                qsc = __ldg(reinterpret_cast<const float4*>(d.q_scale + gc));
                qsh = __ldg(reinterpret_cast<const float4*>(d.q_shift + gc));
            }
            // pixel -> (image, row, col) of the gathered operand, kept incrementally: k-blocks are visited in order and
            // each slot advances by 32 pixels per block (no divisions in the loop)
            struct Pix { int b, gi, gj; };
            auto decode = [&](long long m) {
                Pix x; x.b = (int)(m / p.HgWg); const int r = (int)(m - (long long)x.b * p.HgWg);
                x.gi = r / d.Wg; x.gj = r - x.gi * d.Wg; return x;
            };
            auto advance = [&](Pix& x) {
                x.gj += Cfg::KB;
                while (x.gj >= d.Wg) { x.gj -= d.Wg; if (++x.gi == d.Hg) { x.gi = 0; ++x.b; } }
            };
            constexpr int GSL = RIT > SIT ? RIT : SIT;
            Pix gpx[GSL];
#pragma unroll
            for (int i = 0; i < GSL; ++i) gpx[i] = decode(p.swap ? (mbeg + (s_active ? sp0 : 0) + SPIX * i) : (mbeg + rp0 + RSTEP * i));
            auto fetch = [&](const OpSrc& o, long long m, const Pix& x, int c, bool cvalid) -> float4 {
                float4 v = make_float4(__int_as_float(0x7fc00001), 0.f, 0.f, 0.f);       // "stays zero"
                if (!cvalid || m >= mend) return v;
                if (o.gather) {
                    const int iy = x.gi * d.sy + gky + d.offy, ix = x.gj * d.sx + gkx + d.offx;
                    if (iy < 0 || iy >= d.Hq || ix < 0 || ix >= d.Wq) return v;
                    return __ldg(reinterpret_cast<const float4*>(o.ptr + (((x.b * d.Hq + iy) * d.Wq + ix) * o.ld + c)));
                }
                return __ldg(reinterpret_cast<const float4*>(o.ptr + ((int)m * o.ld + c)));
            };
            auto put = [&](uint8_t* hi_img, uint8_t* lo_img, uint32_t off, float4 v, bool gathered) {
                if (__float_as_int(v.x) == 0x7fc00001) v = make_float4(0.f, 0.f, 0.f, 0.f);
                else if (gathered && d.q_scale) {
                    v.x = fmaf(v.x, qsc.x, qsh.x); v.y = fmaf(v.y, qsc.y, qsh.y); v.z = fmaf(v.z, qsc.z, qsh.z); v.w = fmaf(v.w, qsc.w, qsh.w);
                    if (d.q_relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                }
                float4 hi = split_hi4(v);
                *reinterpret_cast<float4*>(hi_img + off) = hi;
                if (!one) *reinterpret_cast<float4*>(lo_img + off) = split_lo4(v, hi);
            };
            auto load_block = [&](int kb, float4 (&vr)[RIT], float4 (&vs)[SIT]) {
                if (kb >= nkb) return;
                const long long mb = mbeg + (long long)kb * Cfg::KB;
#pragma unroll
                for (int i = 0; i < RIT; ++i) vr[i] = fetch(R, mb + rp0 + RSTEP * i, gpx[i], rc, rcv);
#pragma unroll
                for (int i = 0; i < SIT; ++i) vs[i] = fetch(S, mb + sp0 + SPIX * i, gpx[i], sc_, scv);
#pragma unroll
                for (int i = 0; i < GSL; ++i) advance(gpx[i]);
            };
            auto store_block = [&](int kb, const float4 (&vr)[RIT], const float4 (&vs)[SIT]) {
                if (kb >= nkb) return;
                const int s = kb % NSTAGE; const uint32_t ph = (kb / NSTAGE) & 1;
                mbar_wait(empty(s), ph ^ 1u);
                uint8_t* r_hi = sgen + s * Cfg::STAGE;
                uint8_t* r_lo = r_hi + Cfg::R_BYTES;
                uint8_t* s_hi = r_lo + Cfg::R_BYTES;
                uint8_t* s_lo = s_hi + Cfg::S_BYTES;
#pragma unroll
                for (int i = 0; i < RIT; ++i) put(r_hi, r_lo, mn_off(rp0 + RSTEP * i, rch, 4), vr[i], R.gather != 0);
                if (s_active) {
#pragma unroll
                    for (int i = 0; i < SIT; ++i) put(s_hi, s_lo, mn_off(sp0 + SPIX * i, sch, BN / 32), vs[i], S.gather != 0);
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(full(s));
            };
            // register ring, two k-blocks of gathers in flight while a third is transformed and stored
            float4 r0[RIT], r1[RIT], r2[RIT], s0[SIT], s1[SIT], s2[SIT];
            load_block(0, r0, s0);
            load_block(1, r1, s1);
            for (int kb = 0; kb < nkb; kb += 3) {
                load_block(kb + 2, r2, s2); store_block(kb, r0, s0);
                load_block(kb + 3, r0, s0); store_block(kb + 1, r1, s1);
                load_block(kb + 4, r1, s1); store_block(kb + 2, r2, s2);
            }
            // ---------------- epilogue: TMEM -> fp32 atomics into dw ----------------
            mbar_wait(accum_bar, 0u);
            tc_fence_after();
            const int q = warp & 3, cgrp = warp >> 2;        // TMEM lane quarter, column group (16 warps -> 4 groups)
            const int row = q * 32 + lane;                   // M-side column within the tile
            const int mcol = R.c0 + row;
            const int nacc = nkb < NACC ? nkb : NACC;
            for (int c0 = cgrp * 16; c0 < BN; c0 += 64) {
                if (S.c0 + c0 >= S.C) break;
                float v[16];
                tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
                for (int a = 1; a < nacc; ++a) {
                    float u[16];
                    tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * BN + c0), u);
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] += u[j];
                }
                if (mcol < R.C) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int ncol = S.c0 + c0 + j;
                        if (ncol < S.C) {
                            const int kq = p.swap ? ncol : mcol, ca = p.swap ? mcol : ncol;
                            atomicAdd(d.dw + (size_t)kq * d.Ca + ca, v[j]);
                        }
                    }
                }
            }
            tc_fence_before();
        }
    } else {
        // ---------------- MMA issuer: converged warp, one elected lane issues (see conv_halo_tma.cu) ----------------
        const bool leader = elect_one();
        if (nkb > 0) {
            // D=f32, A=B=tf32, both MN-major (bits 15, 16), N=BN, M=128
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                                   ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % NSTAGE; const uint32_t ph = (kb / NSTAGE) & 1;
                mbar_wait(full(s), ph);
                tc_fence_after();
                const uint32_t r_hi = sbase + s * Cfg::STAGE;
                const uint32_t r_lo = r_hi + Cfg::R_BYTES;
                const uint32_t s_hi = r_lo + Cfg::R_BYTES;
                const uint32_t s_lo = s_hi + Cfg::S_BYTES;
                const uint32_t acc = tmem + (uint32_t)((kb % NACC) * BN);
                const uint32_t fresh = (kb < NACC) ? 0u : 1u;
                if (leader) {
                const uint64_t tR = make_desc_mn(0, 512, 2048), tS = make_desc_mn(0, 512, BN * 16);      // templates: the address moves the low 14 bits only
                const uint64_t drh0 = tR + (uint64_t)(r_hi >> 4), drl0 = tR + (uint64_t)(r_lo >> 4);
                const uint64_t dsh0 = tS + (uint64_t)(s_hi >> 4), dsl0 = tS + (uint64_t)(s_lo >> 4);
#pragma unroll
                for (int j = 0; j < 4; ++j) {                 // 4 groups of 8 pixels
                    const uint64_t drh = drh0 + (uint64_t)(j * 4096 >> 4), drl = drl0 + (uint64_t)(j * 4096 >> 4);
                    const uint64_t dsh = dsh0 + (uint64_t)(j * (BN * 32) >> 4), dsl = dsl0 + (uint64_t)(j * (BN * 32) >> 4);
                    if (!one) {
                        mma_tf32(acc, drl, dsh, idesc, (j ? 1u : fresh));
                        mma_tf32(acc, drh, dsl, idesc, 1u);
                    }
                    mma_tf32(acc, drh, dsh, idesc, one ? (j ? 1u : fresh) : 1u);
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

template <int BN>
static int launch_wg(const WgTcP& p, dim3 grid, cudaStream_t st) {
    using Cfg = WgCfg<BN>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_wgrad_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
        if (e != cudaSuccess) { set_error("conv_wgrad_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); return SAUNET_ERR_CUDA; }
        attr_set = true;
    }
    conv_wgrad_tc_kernel<BN><<<grid, kWgThreads, Cfg::SMEM, st>>>(p);
    SAUNET_CHECK_LAUNCH("conv_wgrad_tc_kernel");
    return SAUNET_OK;
}

bool conv_wgrad_tc_eligible(const saunet_wgrad_desc* d) {
    if (d->precision != 1 && d->precision != 2) return false;
    if (d->Ca % 4 || d->Cb % 4 || d->p_ld % 4 || d->q_ld % 4 || !aligned16(d->p) || !aligned16(d->q)) return false;
    if (d->q_scale && (!aligned16(d->q_scale) || !aligned16(d->q_shift))) return false;
    if (d->Ca < 8 || d->Cb < 4 || d->KH * d->KW * d->Cb < 32) return false;     // (Cb = 4: the channel-padded 7x7 stem)
    const long long M = (long long)d->B * d->Hg * d->Wg;                       // 32-bit element offsets in the kernel
    if (M * d->p_ld >= (1ll << 31) || (long long)d->B * d->Hq * d->Wq * d->q_ld >= (1ll << 31)) return false;
    return true;
}

int conv_wgrad_tc(const saunet_wgrad_desc* d, cudaStream_t st) {
    WgTcP p; p.d = *d;
    p.M = (long long)d->B * d->Hg * d->Wg; p.HgWg = d->Hg * d->Wg;
    SAUNET_CHECK_ARG(p.M > 0, SAUNET_ERR_BAD_SHAPE, "conv2d_wgrad(tc): empty problem");
    // the wider operand goes on the 128-row M side; the gathered operand's width is its im2col width taps*Cb
    const int CQ = d->KH * d->KW * d->Cb;
    p.swap = (d->Ca > CQ) ? 1 : 0;
    const int Cm = p.swap ? d->Ca : CQ, Cn = p.swap ? CQ : d->Ca;
    const int BN = Cn <= 32 ? 32 : (Cn <= 64 ? 64 : 128);
    p.mtiles = cdiv(Cm, 128); p.ntiles = cdiv(Cn, BN);
    const long long tiles = (long long)p.mtiles * p.ntiles;
    SAUNET_CHECK_ARG(tiles <= 65535, SAUNET_ERR_BAD_SHAPE, "conv2d_wgrad(tc): too many tiles");
    // split the pixel axis so that ~2 CTAs per SM exist; at least 256 pixels per split
    long long splits = ((long long)kNumSMs * 2 + tiles - 1) / tiles;
    long long maxs = (p.M + 255) / 256;
    if (splits > maxs) splits = maxs;
    if (splits < 1) splits = 1;
    long long pps = (p.M + splits - 1) / splits;
    pps = (pps + 31) / 32 * 32;
    splits = (p.M + pps - 1) / pps;
    p.pix_per_split = pps;
    dim3 grid((unsigned)splits, (unsigned)tiles);
    if (BN == 32) return launch_wg<32>(p, grid, st);
    if (BN == 64) return launch_wg<64>(p, grid, st);
    return launch_wg<128>(p, grid, st);
}

}  // namespace saunet
