// conv_halo_tma.cu -- 3x3 / stride 1 / pad 1 convolution (forward and data-gradient) on tcgen05, fed by TMA.
//
// Same tile algebra as conv_halo_persist.cu (16 x 8-pixel tile = the 128 TMEM lanes, 18 x 10 halo patch of one
// 32-channel chunk in the UMMA K-major SWIZZLE_128B pattern, the 9 taps = 9 shifted A descriptors over ONE patch), but
// the patch is no longer gathered by LDG through registers:
//   * one thread issues `cp.async.bulk.tensor.4d` (TMA, tile mode) over the NHWC activation viewed as [B][H][W][C]:
//     box = {32 ch, 10, 18, 1} lands as 180 dense 128-byte rows, swizzled by the TMA unit, zero-filled outside the image
//     (= the convolution's zero padding) and beyond the last channel (Cin = 16 / 48 layers get their K padding for free);
//   * 8 transform warps turn the raw fp32 patch into MMA operands IN PLACE (shared -> registers -> shared): fused
//     BN + ReLU prologue (re-zeroing the padding, which the prologue must not touch) and the tf32 split -- with the
//     truncation split the raw value IS the hi operand, so a prologue-free operand (every data gradient) only gets
//     its lo image written;
//   * three patch buffers (two for 128-wide tiles) keep two TMA loads in flight while the MMAs of a third run, across
//     tile boundaries (persistent, one CTA per SM), so neither the gather latency nor the per-tile setup that bound
//     the narrow tiles of conv_halo.cu (2 CTAs / SM, one tile each, 32-44 % tensor-pipe active) is exposed.
// Warps: 0-7 transform, 8-15 epilogue (second TMEM accumulator buffer), 16 MMA issue + TMEM alloc, 17 weight-tile
// ring (cp.async.bulk), 18 patch TMA.
#include "tc_common.cuh"
#include <cuda.h>
#include <type_traits>

namespace saunet {

struct HaloTP {
    saunet_conv_desc d;
    int tiles_x, tiles_y, nchunk;
    int ntile_n, ntiles;                          // tile id = spatial tile * ntile_n + n tile
    const float* wt;
    long long* prof;                              // optional per-CTA cycle counters (tools/bench_conv.py, SAUNET_TC_PROF): [148][16]
};
#define TP_T0() long long pt0__ = p.prof ? clock64() : 0
#define TP_ADD(var) do { if (p.prof) { long long t1__ = clock64(); (var) += t1__ - pt0__; pt0__ = t1__; } } while (0)

constexpr int kTPitch = 10;                         // patch rows per image row: stride-byte-offset 1280 (verified: any multiple of 128 works)
constexpr int kTPatchRows = 18 * kTPitch;
constexpr int kTPatchBytes = (kTPatchRows * 128 + 1023) / 1024 * 1024;   // one image (hi or lo); 1024-aligned so that
                                                                       // 'row index & 7' IS the absolute-address swizzle phase
constexpr int kTHaloItems = 18 * 10 * 8;            // 16-byte chunks per patch
constexpr int kHaloTXform = 256;
constexpr int kTHaloEpilogue = 256;
constexpr int kTHaloThreads = kHaloTXform + kTHaloEpilogue + 96;
constexpr uint32_t kTBoxBytes = 18 * 10 * 128;                 // bytes one TMA box delivers (always the full box; out of bounds = zeros)
constexpr int kTHaloIters = (kTHaloItems + kHaloTXform - 1) / kHaloTXform;   // 6

template <int BN, int NPASS>
struct HaloTCfg {
    using Op = Opnd<NPASS>;
    static constexpr int NOP = Op::NOP;
    // patch buffers (tf32: wide tiles need the room for weight stages).  bf16: the MMAs of a chunk take ~600 clocks while a
    // TMA box (180 separate 128-byte rows) takes several thousand to land -- ncu on 64 -> 64 at 256x256 with two buffers:
    // 13 % tensor-active, 21 % DRAM, nothing saturated -- so four boxes are kept in flight
#ifndef SAUNET_TMA_BF16_NBUF
#define SAUNET_TMA_BF16_NBUF 4
#endif
    static constexpr int NBUF = Op::BF ? SAUNET_TMA_BF16_NBUF : (BN >= 64 ? 2 : 3);
    // bf16 operands: the raw fp32 box is followed by the bf16 operand image (64-byte rows, SWIZZLE_64B) the transform warps
    // write and the MMAs read; 12 KB keeps every buffer 1024-byte aligned (what the TMA swizzle needs)
    static constexpr int BF_IMG = 12 * 1024;
    static constexpr int PATCH = Op::BF ? kTPatchBytes + BF_IMG : NOP * kTPatchBytes;      // one buffer
    static constexpr int B_TAP = NOP * BN * Op::ROW;                  // weight image of one (chunk, tap)
    // Narrow tiles issue only 8 short MMAs per tap: one barrier round trip per tap (~100 clocks of try_wait + commit on
    // the single issuing thread) would cost as much as the MMAs themselves, so a weight stage holds G consecutive taps.
#ifndef SAUNET_TMA_BF16_G
#define SAUNET_TMA_BF16_G(bn) ((bn) <= 64 ? 9 : 3)
#endif
    static constexpr int G = Op::BF ? SAUNET_TMA_BF16_G(BN) : (BN <= 16 ? 9 : (BN <= 64 ? 3 : 1));   // (bf16: 2 MMAs per tap)
    static constexpr int B_STAGE = G * B_TAP;
    static constexpr int RED_BYTES = 8 * BN * 4;
    static constexpr int B_SPACE = 224 * 1024 - NBUF * PATCH - RED_BYTES;
    static constexpr int NSTB_RAW = B_SPACE / B_STAGE;
    static constexpr int NSTB = NSTB_RAW > 4 ? 4 : NSTB_RAW;
    static constexpr int SMEM = NBUF * PATCH + NSTB * B_STAGE + RED_BYTES + 1024 + 256;
    // Narrow tiles are bound by the tensor core's operand reads from shared memory (the 128x32 fp32 A tile of every MMA
    // is 4 KB whatever N is), not by its math: for BN <= 64 the hi and lo weight images, adjacent in the stage, are fed
    // as ONE N = 2*BN operand, so A_hi is read once for the hi*hi and hi*lo products (2 MMAs per K step instead of 3);
    // the two halves accumulate in separate TMEM columns and are added in the epilogue.
    static constexpr bool CAT = (NPASS == 3) && (BN <= 64);
    static constexpr int ACC_COLS = (BN < 32 ? 32 : BN) * (CAT ? 2 : 1);
    // (bf16 operands: one accumulator -- the round-robin over several only serves the fp32-class error budget of 3xTF32, and
    //  every extra accumulator is one more TMEM load per 16 columns in an epilogue that bounds the short-K tiles)
    static constexpr int NACC = Op::BF ? 1 : ((256 / ACC_COLS) > 4 ? 4 : (256 / ACC_COLS));
    static constexpr int BUF_COLS = NACC * ACC_COLS;
    static constexpr int TMEM_COLS = 2 * BUF_COLS <= 256 ? 256 : 512;
    static_assert(NSTB >= 2, "weight ring needs two stages");
};

template <int BN, int NPASS>
__global__ void __launch_bounds__(kTHaloThreads, 1) conv_halo_tma_kernel(const __grid_constant__ HaloTP p, const __grid_constant__ CUtensorMap tmap) {
    using Cfg = HaloTCfg<BN, NPASS>;
    constexpr int NSTB = Cfg::NSTB, NACC = Cfg::NACC;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
    const uint32_t b_base = sbase + Cfg::NBUF * Cfg::PATCH;
    float* red = reinterpret_cast<float*>(sgen + Cfg::NBUF * Cfg::PATCH + NSTB * Cfg::B_STAGE);
    const uint32_t bars = b_base + NSTB * Cfg::B_STAGE + Cfg::RED_BYTES;
    constexpr int NBUF = Cfg::NBUF;
    auto raw_full = [&](int i) { return bars + 8u * i; };                   // TMA box landed (complete_tx)
    auto patch_full = [&](int i) { return bars + 8u * (NBUF + i); };        // transform warps done: operands ready
    auto patch_empty = [&](int i) { return bars + 8u * (2 * NBUF + i); };   // the MMAs that read the buffer retired
    auto tmem_full = [&](int i) { return bars + 8u * (3 * NBUF + i); };
    auto tmem_empty = [&](int i) { return bars + 8u * (3 * NBUF + 2 + i); };
    auto b_full = [&](int s) { return bars + 8u * (3 * NBUF + 4 + s); };
    auto b_empty = [&](int s) { return bars + 8u * (3 * NBUF + 4 + NSTB + s); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sgen + Cfg::NBUF * Cfg::PATCH + NSTB * Cfg::B_STAGE + Cfg::RED_BYTES + 8 * (3 * NBUF + 4 + 2 * NSTB));

    const saunet_conv_desc& d = p.d;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int NPW = kHaloTXform / 32, EPI_WARP0 = NPW, MMA_WARP = NPW + kTHaloEpilogue / 32, LOAD_WARP = MMA_WARP + 1, TMA_WARP = MMA_WARP + 2;
    const int my_tiles = (p.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;    // >= 1
    const int nchunk = p.nchunk;
    auto tile_coords = [&](int ti, int& b, int& y0, int& x0, int& n0) {
        int t = (int)blockIdx.x + ti * (int)gridDim.x;
        n0 = (t % p.ntile_n) * BN; t /= p.ntile_n;
        const int txi = t % p.tiles_x; t /= p.tiles_x;
        const int tyi = t % p.tiles_y; b = t / p.tiles_y;
        y0 = tyi * 16; x0 = txi * 8;
    };

    if (tid == 0) {
        for (int i = 0; i < NBUF; ++i) { mbar_init(raw_full(i), 1); mbar_init(patch_full(i), kHaloTXform / 32); mbar_init(patch_empty(i), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(tmem_full(i), 1); mbar_init(tmem_empty(i), kTHaloEpilogue / 32); }
        for (int s = 0; s < NSTB; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
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

    if (warp == TMA_WARP) {
        // ================= patch loads: one TMA box per (tile, 32-channel chunk) =================
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap)) : "memory");
            int f = 0;
            long long w_pe = 0;
            const long long kt0 = p.prof ? clock64() : 0;
            for (int ti = 0; ti < my_tiles; ++ti) {
                int b, y0, x0, n0; tile_coords(ti, b, y0, x0, n0);
                for (int cc = 0; cc < nchunk; ++cc, ++f) {
                    const int buf = f % NBUF; const uint32_t ph = (f / NBUF) & 1;
                    TP_T0();
                    mbar_wait(patch_empty(buf), ph ^ 1u);
                    TP_ADD(w_pe);
                    mbar_expect_tx(raw_full(buf), kTBoxBytes);
                    // coordinates innermost first: channel, x, y, image; negative / past-the-end = zero fill
                    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                                 ::"r"(sbase + (uint32_t)(buf * Cfg::PATCH)), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(cc * 32), "r"(x0 - 1),
                                   "r"(y0 - 1), "r"(b), "r"(raw_full(buf)) : "memory");
                }
            }
            if (p.prof) { long long* o = p.prof + blockIdx.x * 16; o[0] = clock64() - kt0; o[1] = w_pe; }
        }
        __syncwarp();
    } else if (warp < NPW) {
        // ================= transform warps: raw fp32 patch -> MMA operands, in place =================
        const int chunk = tid & 7;
        uint32_t s_off[kTHaloIters];      // byte offset inside a patch image (tile independent)
        uint32_t b_off[kTHaloIters];      // bf16 operands: byte offset inside the bf16 image
        int p_rc[kTHaloIters];            // patch (row << 8 | col), -1: no such item
#pragma unroll
        for (int i = 0; i < kTHaloIters; ++i) {
            const int it = tid + kHaloTXform * i;
            const int pix = it >> 3;
            const int py = pix / 10, px = pix - py * 10;
            const int pr = py * kTPitch + px;
            p_rc[i] = it < kTHaloItems ? ((py << 8) | px) : -1;
            s_off[i] = (uint32_t)pr * 128u + (uint32_t)((chunk ^ (pr & 7)) << 4);
            b_off[i] = Opnd<kBF16>::off(pr, chunk);
        }
        const bool pro = d.in_scale != nullptr;
        int f = 0;
        long long x_wait = 0, x_work = 0;
        for (int ti = 0; ti < my_tiles; ++ti) {
            unsigned inb = 0;             // items that are real pixels (the prologue must leave the zero padding alone)
            if (pro) {
                int b, y0, x0, n0; tile_coords(ti, b, y0, x0, n0);
#pragma unroll
                for (int i = 0; i < kTHaloIters; ++i) {
                    const int iy = y0 - 1 + (p_rc[i] >> 8), ix = x0 - 1 + (p_rc[i] & 255);
                    if (p_rc[i] >= 0 && iy >= 0 && iy < d.Hin && ix >= 0 && ix < d.Win) inb |= 1u << i;
                }
            }
            for (int cc = 0; cc < nchunk; ++cc, ++f) {
                const int buf = f % NBUF; const uint32_t ph = (f / NBUF) & 1;
                float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
                const int c = cc * 32 + chunk * 4;
                const bool cval = c < d.Cin;                       // channels past Cin (K padding) stay zero
                if (pro && cval) {
                    sc = __ldg(reinterpret_cast<const float4*>(d.in_scale + c));
                    sh = __ldg(reinterpret_cast<const float4*>(d.in_shift + c));
                }
                uint8_t* hi_img = sgen + buf * Cfg::PATCH;
                uint8_t* lo_img = hi_img + kTPatchBytes;
                TP_T0();
                mbar_wait(raw_full(buf), ph);
                TP_ADD(x_wait);
                if (pro || NPASS == 3 || Opnd<NPASS>::BF) {
                    float4 v[kTHaloIters];
#pragma unroll
                    for (int i = 0; i < kTHaloIters; ++i)
                        if (p_rc[i] >= 0) v[i] = *reinterpret_cast<const float4*>(hi_img + s_off[i]);
#pragma unroll
                    for (int i = 0; i < kTHaloIters; ++i) {
                        if (p_rc[i] < 0) continue;
                        float4 tv = v[i];
                        if (pro) {
                            if ((inb & (1u << i)) && cval) {
                                tv.x = fmaf(tv.x, sc.x, sh.x); tv.y = fmaf(tv.y, sc.y, sh.y); tv.z = fmaf(tv.z, sc.z, sh.z); tv.w = fmaf(tv.w, sc.w, sh.w);
                                if (d.in_relu) { tv.x = fmaxf(tv.x, 0.f); tv.y = fmaxf(tv.y, 0.f); tv.z = fmaxf(tv.z, 0.f); tv.w = fmaxf(tv.w, 0.f); }
                            }
                            if (!Opnd<NPASS>::BF) *reinterpret_cast<float4*>(hi_img + s_off[i]) = split_hi4(tv);
                        }
                        if (Opnd<NPASS>::BF) { Opnd<kBF16>::store(lo_img, nullptr, b_off[i], tv); continue; }     // raw fp32 -> bf16 image
#ifdef SAUNET_SPLIT_RN
                        else if (NPASS == 3) *reinterpret_cast<float4*>(hi_img + s_off[i]) = split_hi4(tv);
#endif
                        if (NPASS == 3) *reinterpret_cast<float4*>(lo_img + s_off[i]) = split_lo4(tv, split_hi4(tv));
                    }
                    fence_proxy_async();
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(patch_full(buf));
                TP_ADD(x_work);
            }
        }
        if (p.prof && tid == 0) { long long* o = p.prof + blockIdx.x * 16; o[2] = x_wait; o[3] = x_work; }
    } else if (warp == MMA_WARP) {
        // The whole warp runs the loop CONVERGED (all lanes poll the barriers, all lanes hold the same uniform state) and
        // one elected lane issues the MMAs / commits.  With the loop inside `if (lane == 0)` the compiler wraps every
        // tcgen05.mma in an ELECT / BRA.U.ANY loop and rebuilds both 64-bit descriptors from scratch -- ~17 dependent
        // uniform-datapath instructions, 70 (tf32) to 170 (bf16) clocks per MMA on the one issuing thread, which the
        // per-role counters showed to be THE bound of every narrow tile (mma:issue 76-97 % of the kernel).  Descriptors are
        // now templates (address field 0) advanced by 32-bit adds: an operand address only moves the low 14 bits.
        using Op = Opnd<NPASS>;
        const uint32_t idesc = Op::idesc(BN);
        const uint32_t idesc2 = Op::idesc(2 * BN);
        const uint64_t a_tmpl = Op::desc(0, kTPitch * Op::ROW), b_tmpl = Op::desc(0);
        const bool leader = elect_one();
        int f = 0, g = 0;                     // flat patch / weight-stage counters
        long long m_te = 0, m_pf = 0, m_bf = 0, m_is = 0;
        for (int ti = 0; ti < my_tiles; ++ti) {
            const int abuf = ti & 1; const uint32_t tph = (ti >> 1) & 1;
            TP_T0();
            mbar_wait(tmem_empty(abuf), tph ^ 1u);
            TP_ADD(m_te);
            tc_fence_after();
            int kb = 0;
            for (int cc = 0; cc < nchunk; ++cc, ++f) {
                const int buf = f % NBUF; const uint32_t pph = (f / NBUF) & 1;
                mbar_wait(patch_full(buf), pph);
                TP_ADD(m_pf);
                const int krem = d.Cin - cc * 32;                                                 // zero-padded K tail: skip it
                const int ksteps = krem >= 32 ? Op::KSTEPS : (Op::BF ? (krem + 15) / 16 : (krem + 7) / 8);
                const uint32_t a_hi0 = sbase + buf * Cfg::PATCH;
                const uint64_t a_hi_d = a_tmpl + (uint64_t)(a_hi0 >> 4), a_lo_d = a_hi_d + (uint64_t)(kTPatchBytes >> 4);
                for (int tg = 0; tg < 9 / Cfg::G; ++tg, ++g) {
                    const int s = g % NSTB; const uint32_t ph = (g / NSTB) & 1;
                    mbar_wait(b_full(s), ph);
                    TP_ADD(m_bf);
                    tc_fence_after();
                    if (leader) {
                        const uint64_t b_d0 = b_tmpl + (uint64_t)((b_base + s * Cfg::B_STAGE) >> 4);
                        // `nk`: K steps of this chunk -- a compile-time constant for whole chunks (fully unrolled, the MMAs of
                        // a tap group issue back to back), a run-time count only for the zero-padded last chunk
                        auto issue = [&](auto nk) {
#pragma unroll
                            for (int t = 0; t < Cfg::G; ++t) {
                                const int tap = tg * Cfg::G + t;
                                const int ky = tap / 3, kx = tap - ky * 3;
                                const uint64_t shift = (uint64_t)(((ky * kTPitch + kx) * Op::ROW) >> 4);
                                const uint64_t b_hi_d = b_d0 + (uint64_t)((t * Cfg::B_TAP) >> 4), b_lo_d = b_hi_d + (uint64_t)((BN * 128) >> 4);
                                const int kbt = kb + t;
                                const uint32_t acc = tmem + (uint32_t)(abuf * Cfg::BUF_COLS + (kbt % NACC) * Cfg::ACC_COLS);
                                const uint32_t fresh = (kbt < NACC) ? 0u : 1u;
#pragma unroll
                                for (int kk = 0; kk < nk; ++kk) {
                                    if (Op::BF) {          // the operand is the bf16 image behind the raw patch
                                        Op::mma(acc, a_lo_d + shift + (uint64_t)(kk * 2), b_hi_d + (uint64_t)(kk * 2), idesc, (kk ? 1u : fresh));
                                        continue;
                                    }
                                    const uint64_t dah = a_hi_d + shift + (uint64_t)(kk * 2), dbh = b_hi_d + (uint64_t)(kk * 2);
                                    if (Cfg::CAT) {
                                        const uint64_t dal = a_lo_d + shift + (uint64_t)(kk * 2);
                                        mma_tf32(acc, dah, dbh, idesc2, (kk ? 1u : fresh));       // [hi*hi | hi*lo]: B rows BN..2BN-1 are the lo image
                                        mma_tf32(acc + BN, dal, dbh, idesc, 1u);                  // lo*hi joins the small-terms half
                                    } else if (NPASS == 3) {
                                        const uint64_t dal = a_lo_d + shift + (uint64_t)(kk * 2), dbl = b_lo_d + (uint64_t)(kk * 2);
                                        mma_tf32(acc, dal, dbh, idesc, (kk ? 1u : fresh));
                                        mma_tf32(acc, dah, dbl, idesc, 1u);
                                        mma_tf32(acc, dah, dbh, idesc, 1u);
                                    } else {
                                        mma_tf32(acc, dah, dbh, idesc, (kk ? 1u : fresh));
                                    }
                                }
                            }
                        };
                        if (krem >= 32) issue(std::integral_constant<int, Op::KSTEPS>{}); else issue(ksteps);
                        mma_commit(b_empty(s));
                    }
                    kb += Cfg::G;
                    __syncwarp();
                    TP_ADD(m_is);
                }
                if (leader) mma_commit(patch_empty(buf));
            }
            if (leader) mma_commit(tmem_full(abuf));
            __syncwarp();
        }
        if (p.prof && leader) { long long* o = p.prof + blockIdx.x * 16; o[4] = m_te; o[5] = m_pf; o[6] = m_bf; o[7] = m_is; }
        __syncwarp();
    } else if (warp == LOAD_WARP) {
        if (lane == 0) {
            constexpr uint32_t BYTES = Cfg::B_STAGE;
            const int nkb = nchunk * 9 / Cfg::G;          // weight stages per tile (G consecutive taps each)
            int g = 0;
            long long l_be = 0;
            for (int ti = 0; ti < my_tiles; ++ti) {
                const int nt = ((int)blockIdx.x + ti * (int)gridDim.x) % p.ntile_n;
                const uint8_t* src = reinterpret_cast<const uint8_t*>(p.wt) + (size_t)nt * nkb * BYTES;
                for (int kb = 0; kb < nkb; ++kb, ++g) {
                    const int s = g % NSTB; const uint32_t ph = (g / NSTB) & 1;
                    TP_T0();
                    mbar_wait(b_empty(s), ph ^ 1u);
                    TP_ADD(l_be);
                    mbar_expect_tx(b_full(s), BYTES);
                    bulk_g2s(b_base + s * Cfg::B_STAGE, src + (size_t)kb * BYTES, BYTES, b_full(s));
                }
            }
            if (p.prof) { long long* o = p.prof + blockIdx.x * 16; o[8] = l_be; }
        }
        __syncwarp();
    } else {
        // ================= epilogue =================
        const int q = warp & 3, half = (warp - EPI_WARP0) >> 2;
        const int etid = tid - EPI_WARP0 * 32;
        const int row = q * 32 + lane;
        const bool vst = (d.Cout % 4 == 0) && (d.y_ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(d.y) & 15u) == 0);
        const int nkb = nchunk * 9;
        const int nacc = nkb < NACC ? nkb : NACC;
        long long e_wait = 0, e_work = 0;
        for (int ti = 0; ti < my_tiles; ++ti) {
            const int abuf = ti & 1; const uint32_t tph = (ti >> 1) & 1;
            int b, y0, x0, n0; tile_coords(ti, b, y0, x0, n0);
            const int oy = y0 + (row >> 3), ox = x0 + (row & 7);
            const size_t m = (size_t)(b * d.Hout + oy) * d.Wout + ox;
            float* yp = d.y + m * d.y_ld;
            const float rs = d.row_scale ? (d.row_scale[m] + d.row_scale_add) : 1.f;
            if (ti > 0 && d.stat_sum) asm volatile("bar.sync 1, 256;" ::: "memory");      // red[] of the previous tile consumed
            TP_T0();
            mbar_wait(tmem_full(abuf), tph);
            TP_ADD(e_wait);
            tc_fence_after();
            const uint32_t tb = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(abuf * Cfg::BUF_COLS);
            for (int c0 = half * 16; c0 < BN; c0 += 32) {
                if (n0 + c0 >= d.Cout) break;
                float v[16];
                tmem_ld16(tb + (uint32_t)c0, v);
                for (int a = 0; a < nacc; ++a) {
                    float u[16];
                    if (a > 0) {
                        tmem_ld16(tb + (uint32_t)(a * Cfg::ACC_COLS + c0), u);
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] += u[j];
                    }
                    if (Cfg::CAT) {                               // the hi*lo + lo*hi half of the accumulator
                        tmem_ld16(tb + (uint32_t)(a * Cfg::ACC_COLS + BN + c0), u);
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] += u[j];
                    }
                }
                float o[16];
                epi_chunk(v, o, d.bias ? d.bias + n0 + c0 : nullptr, d.Cout - (n0 + c0), true, true, d.row_scale != nullptr, rs, d.act);
                if (vst) {
#pragma unroll
                    for (int qq = 0; qq < 4; ++qq) {
                        const int n = n0 + c0 + 4 * qq;
                        if (n < d.Cout) {
                            float4* dst = reinterpret_cast<float4*>(yp + n);
                            float4 w4 = make_float4(o[4 * qq], o[4 * qq + 1], o[4 * qq + 2], o[4 * qq + 3]);
                            if (d.accumulate) { float4 cur = *dst; w4.x += cur.x; w4.y += cur.y; w4.z += cur.z; w4.w += cur.w; }
                            *dst = w4;
                        }
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int n = n0 + c0 + j;
                        if (n < d.Cout) yp[n] = d.accumulate ? yp[n] + o[j] : o[j];
                    }
                }
                if (d.stat_sum) {
                    float sq[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) sq[j] = v[j] * v[j];
                    const float s1 = colsum16(v, lane);
                    const float s2 = colsum16(sq, lane);
                    if ((lane & 1) == 0) {
                        red[(q * 2 + 0) * BN + c0 + (lane >> 1)] = s1;
                        red[(q * 2 + 1) * BN + c0 + (lane >> 1)] = s2;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty(abuf));
            if (d.stat_sum) {
                asm volatile("bar.sync 1, 256;" ::: "memory");
                for (int i = etid; i < BN; i += kTHaloEpilogue) {
                    if (n0 + i < d.Cout) {
                        float s1 = 0.f, s2 = 0.f;
#pragma unroll
                        for (int w = 0; w < 4; ++w) { s1 += red[(w * 2 + 0) * BN + i]; s2 += red[(w * 2 + 1) * BN + i]; }
                        atomicAdd(d.stat_sum + n0 + i, (double)s1);
                        atomicAdd(d.stat_sumsq + n0 + i, (double)s2);
                    }
                }
            }
            TP_ADD(e_work);
        }
        if (p.prof && etid == 0) { long long* o = p.prof + blockIdx.x * 16; o[9] = e_wait; o[10] = e_work; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(Cfg::TMEM_COLS) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// cuTensorMapEncodeTiled through the runtime's driver-entry-point query: no link-time dependency on libcuda, so the
// library still loads (and exports its symbols) on a machine without a driver.
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// the activation slice x[.., c0 : c0+Cin] of an NHWC buffer as a rank-4 tensor (C, W, H, B), box = one halo patch of a
// 32-channel chunk; SWIZZLE_128B so the box lands in the UMMA K-major pattern
static int make_patch_map(const saunet_conv_desc* d, CUtensorMap* tm) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) { set_error("conv_halo_tma: cuTensorMapEncodeTiled is not available from this driver"); return SAUNET_ERR_CUDA; }
    const cuuint64_t gdim[4] = {(cuuint64_t)d->Cin, (cuuint64_t)d->Win, (cuuint64_t)d->Hin, (cuuint64_t)d->B};
    const cuuint64_t gstr[3] = {(cuuint64_t)d->x_ld * 4, (cuuint64_t)d->Win * d->x_ld * 4, (cuuint64_t)d->Hin * d->Win * d->x_ld * 4};
    const cuuint32_t box[4] = {32, 10, 18, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(d->x), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("conv_halo_tma: cuTensorMapEncodeTiled failed (%d)", (int)r); return SAUNET_ERR_CUDA; }
    return SAUNET_OK;
}

template <int BN, int NPASS>
static int launch_halo_tma(const HaloTP& p0, cudaStream_t st) {
    using Cfg = HaloTCfg<BN, NPASS>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_halo_tma_kernel<BN, NPASS>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
        if (e != cudaSuccess) { set_error("conv_halo_tma: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); return SAUNET_ERR_CUDA; }
        attr_set = true;
    }
    HaloTP p = p0;
    CUtensorMap tm;
    const int rc = make_patch_map(&p.d, &tm);
    if (rc != SAUNET_OK) return rc;
    p.ntile_n = cdiv(p.d.Cout, BN);
    p.ntiles = p.d.B * p.tiles_y * p.tiles_x * p.ntile_n;
    static long long* const prof_ptr = []() -> long long* {          // debugging aid, read once per process
        const char* pe = getenv("SAUNET_TC_PROF");
        return pe ? reinterpret_cast<long long*>(strtoull(pe, nullptr, 0)) : nullptr;
    }();
    p.prof = prof_ptr;
    const int grid = p.ntiles < kNumSMs ? p.ntiles : kNumSMs;
    conv_halo_tma_kernel<BN, NPASS><<<grid, kTHaloThreads, Cfg::SMEM, st>>>(p, tm);
    SAUNET_CHECK_LAUNCH("conv_halo_tma_kernel");
    return SAUNET_OK;
}

bool conv_tc_eligible(const saunet_conv_desc* d);

// 3x3 / s1 / p1 on whole 16 x 8 tiles; any Cin % 4 == 0 (a partial last chunk is zero-filled by the TMA unit and needs
// the chunk-major padded weight image, desc.tc_cm); weights tiled 16..128 wide
bool conv_halo_tma_eligible(const saunet_conv_desc* d) {
    if (SAUNET_ENV_FLAG("SAUNET_NO_HALO_TMA")) return false;
    if (!conv_tc_eligible(d)) return false;
    if (d->KH != 3 || d->KW != 3 || d->sy != 1 || d->sx != 1 || d->offy != -1 || d->offx != -1) return false;
    if (d->osy != 1 || d->osx != 1 || d->oy0 != 0 || d->ox0 != 0) return false;
    if (d->Hg != d->Hin || d->Wg != d->Win || d->Hout != d->Hin || d->Wout != d->Win) return false;
    if (d->Hin % 16 || d->Win % 8) return false;
    if (d->Cin % 32 && !d->tc_cm) return false;
    if (d->tc_bn > 128) return false;
    if (d->tc_cm) return true;                   // (the LDG-gather kernels need whole 32-channel chunks)
    // Measured on B200 (tools/dbg/halo_sweep.sh, profiles/r02_halo_sweep.txt).  Narrow tiles are bound by shared-memory
    // bandwidth (the MMAs re-read the 4 KB A tile for every 128 x N x 8 block of math), and staging the raw patch through
    // shared memory costs 23 KB more per chunk than registers do: with a BN+ReLU prologue (hi AND lo rewritten) the
    // register-gather kernel stays ahead while it has >= 2 waves of tiles per SM (114 vs 98 TFLOP/s at 128x128,
    // 128 -> 32); prologue-free operands (only lo written) and small problems (latency-bound: 64 vs 54 at 32x32) are faster
    // through TMA (64 -> 64 at 256x256: 194 vs 158 TFLOP/s).
    // bf16 operands: the MMAs read a quarter of the bytes, the raw patch no longer competes with them for shared-memory
    // bandwidth -> every eligible layer (SAUNET_BF16_HALO_LDG: apply the exclusions measured for tf32 operands)
    if (d->tc_passes == kBF16 && !SAUNET_ENV_FLAG("SAUNET_BF16_HALO_LDG")) return true;
    const long long tiles = (long long)d->B * (d->Hin / 16) * (d->Win / 8);
    if (d->in_scale && d->tc_bn <= 64 && tiles >= 2 * 2 * kNumSMs) return false;
    if (!d->in_scale && d->tc_bn <= 64 && d->Cin >= 256) return false;
    return true;
}

int conv_fwd_halo_tma(const saunet_conv_desc* d, cudaStream_t st) {
    HaloTP p; p.d = *d;
    p.tiles_x = d->Win / 8; p.tiles_y = d->Hin / 16; p.nchunk = (d->Cin + 31) / 32; p.wt = d->w_tc;
    const int np = d->tc_passes;
    switch (d->tc_bn) {
        case 16: return np == kBF16 ? launch_halo_tma<16, kBF16>(p, st) : np != 1 ? launch_halo_tma<16, 3>(p, st) : launch_halo_tma<16, 1>(p, st);
        case 32: return np == kBF16 ? launch_halo_tma<32, kBF16>(p, st) : np != 1 ? launch_halo_tma<32, 3>(p, st) : launch_halo_tma<32, 1>(p, st);
        case 64: return np == kBF16 ? launch_halo_tma<64, kBF16>(p, st) : np != 1 ? launch_halo_tma<64, 3>(p, st) : launch_halo_tma<64, 1>(p, st);
        case 128: return np == kBF16 ? launch_halo_tma<128, kBF16>(p, st) : np != 1 ? launch_halo_tma<128, 3>(p, st) : launch_halo_tma<128, 1>(p, st);
    }
    set_error("conv2d_fwd(halo_tma): unsupported N tile %d", d->tc_bn);
    return SAUNET_ERR_BAD_SHAPE;
}

}  // namespace saunet
