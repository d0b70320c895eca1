// conv_halo_persist.cu -- (wide tiles: BN = 128) 3x3 / stride 1 / pad 1 convolution (forward and data-gradient) on tcgen05 with a shared-memory
// HALO patch: every activation is gathered, BN+ReLU-transformed and tf32-split ONCE per 32-channel chunk and then
// consumed by all 9 taps, instead of 9 times as in the generic implicit-GEMM kernel (conv_tc.cu).
//
// Tile = 16 rows x 8 columns of one image (M = 128) x BN output channels.  Patch = 18 x 10 halo pixels, stored as
// 128-byte rows (32 channels) at a pitch of kPPitch rows per image row, in the UMMA K-major SWIZZLE_128B pattern
// computed from the ABSOLUTE row index.  The A descriptor of tap (ky,kx) simply starts (ky*kPPitch + kx) rows into the
// patch: the hardware applies the 128B swizzle on absolute shared-memory address bits [7,10) (verified on B200 with
// tools/exp/umma_shift_test.cu: shifted starts work with base_offset = 0), and the 8-row groups of the M=128
// operand are the 8-pixel tile rows, kPPitch*128 bytes apart (= the descriptor's stride-byte-offset).
//
// PERSISTENT, warp-specialised (576 threads, one CTA per SM walking tiles; same skeleton as conv_tc.cu):
//   warps 0-7    patch producers: flat loop over (tile, 32-channel chunk), two patch buffers, register prefetch of the
//                next chunk (across tile boundaries)
//   warps 8-15   epilogue on the second TMEM accumulator buffer (bias / BN statistics / row gate / activation / stores)
//   warp 16      TMEM alloc + single-lane MMA issue;   warp 17  weight-tile loader (cp.async.bulk ring; same
//                chunk-major tiled weight image as conv_tc.cu)
// so a tile's prologue latency, its epilogue and the other tiles' MMAs all overlap (the one-tile-per-CTA version ran
// two CTAs per SM for that and still kept the tensor pipe only 32-44 % busy).
#include "tc_common.cuh"

namespace saunet {

struct HaloPP {
    saunet_conv_desc d;
    int tiles_x, tiles_y, nchunk;
    int ntile_n, ntiles;                          // tile id = spatial tile * ntile_n + n tile
    const float* wt;
};

constexpr int kPPitch = 10;                         // patch rows per image row: stride-byte-offset 1280 (verified: any multiple of 128 works)
constexpr int kPPatchRows = 18 * kPPitch;
constexpr int kPPatchBytes = (kPPatchRows * 128 + 1023) / 1024 * 1024;   // one image (hi or lo); 1024-aligned so that
                                                                       // 'row index & 7' IS the absolute-address swizzle phase
constexpr int kPHaloItems = 18 * 10 * 8;            // 16-byte chunks per patch
constexpr int kHaloPProducers = 256;
constexpr int kPHaloEpilogue = 256;
constexpr int kPHaloThreads = kHaloPProducers + kPHaloEpilogue + 64;
constexpr int kPHaloIters = (kPHaloItems + kHaloPProducers - 1) / kHaloPProducers;   // 6

template <int BN, int NPASS>
struct HaloPCfg {
    using Op = Opnd<NPASS>;
    static constexpr int NOP = Op::NOP;
    static constexpr int NBUF = 2;                                   // patch buffers
    static constexpr int IMG = Op::BF ? kPPatchBytes / 2 : kPPatchBytes;   // one patch image (bf16: 64-byte rows)
    static constexpr int PATCH = NOP * IMG;                           // one buffer
    static constexpr int B_STAGE = NOP * BN * Op::ROW;
    static constexpr int RED_BYTES = 8 * BN * 4;
    static constexpr int B_SPACE = 224 * 1024 - NBUF * PATCH - RED_BYTES;
    static constexpr int NSTB_RAW = B_SPACE / B_STAGE;
    static constexpr int NSTB = NSTB_RAW > 6 ? 6 : NSTB_RAW;
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
__global__ void __launch_bounds__(kPHaloThreads, 1) conv_halo_persist_kernel(const __grid_constant__ HaloPP p) {
    using Cfg = HaloPCfg<BN, NPASS>;
    constexpr int NSTB = Cfg::NSTB, NACC = Cfg::NACC;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
    const uint32_t b_base = sbase + Cfg::NBUF * Cfg::PATCH;
    float* red = reinterpret_cast<float*>(sgen + Cfg::NBUF * Cfg::PATCH + NSTB * Cfg::B_STAGE);
    const uint32_t bars = b_base + NSTB * Cfg::B_STAGE + Cfg::RED_BYTES;
    auto patch_full = [&](int i) { return bars + 8u * i; };
    auto patch_empty = [&](int i) { return bars + 8u * (2 + i); };
    auto tmem_full = [&](int i) { return bars + 8u * (4 + i); };
    auto tmem_empty = [&](int i) { return bars + 8u * (6 + i); };
    auto b_full = [&](int s) { return bars + 8u * (8 + s); };
    auto b_empty = [&](int s) { return bars + 8u * (8 + NSTB + s); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sgen + Cfg::NBUF * Cfg::PATCH + NSTB * Cfg::B_STAGE + Cfg::RED_BYTES + 8 * (8 + 2 * NSTB));

    const saunet_conv_desc& d = p.d;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int NPW = kHaloPProducers / 32, EPI_WARP0 = NPW, MMA_WARP = NPW + kPHaloEpilogue / 32, LOAD_WARP = MMA_WARP + 1;
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
        for (int i = 0; i < 2; ++i) {
            mbar_init(patch_full(i), kHaloPProducers / 32); mbar_init(patch_empty(i), 1);
            mbar_init(tmem_full(i), 1); mbar_init(tmem_empty(i), kPHaloEpilogue / 32);
        }
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

    if (warp < NPW) {
        // ================= patch producers: flat loop over (tile, chunk) =================
        const int chunk = tid & 7;
        const int total = my_tiles * nchunk;
        uint32_t s_off[kPHaloIters];      // byte offset inside a patch image (tile independent)
        int p_rc[kPHaloIters];            // patch (row << 8 | col), -1: no such item
#pragma unroll
        for (int i = 0; i < kPHaloIters; ++i) {
            const int it = tid + kHaloPProducers * i;
            const int pix = it >> 3;
            const int py = pix / 10, px = pix - py * 10;
            const int pr = py * kPPitch + px;
            p_rc[i] = it < kPHaloItems ? ((py << 8) | px) : -1;
            s_off[i] = Opnd<NPASS>::off(pr, chunk);
        }
        int g_off[kPHaloIters];           // element offset of the halo pixel (channel chunk*4) in the load cursor's tile, -1 = outside
        int l_ti = -1, l_next_ti = 0, l_cc = 0;
        auto set_tile = [&](int ti) {
            int b, y0, x0, n0; tile_coords(ti, b, y0, x0, n0);
#pragma unroll
            for (int i = 0; i < kPHaloIters; ++i) {
                const int iy = y0 - 1 + (p_rc[i] >> 8), ix = x0 - 1 + (p_rc[i] & 255);
                const bool ok = p_rc[i] >= 0 && iy >= 0 && iy < d.Hin && ix >= 0 && ix < d.Win;
                g_off[i] = ok ? ((b * d.Hin + iy) * d.Win + ix) * d.x_ld + chunk * 4 : -1;
            }
            l_ti = ti;
        };
        // `vm`: validity bits of the loaded items (padding stays exactly zero: the prologue is applied only to real pixels)
        auto load_next = [&](float4 (&v)[kPHaloIters], int& cc_out, unsigned& vm) {
            if (l_next_ti >= my_tiles) return;
            if (l_next_ti != l_ti) set_tile(l_next_ti);
            const int cc = l_cc;
            if (++l_cc == nchunk) { l_cc = 0; ++l_next_ti; }
            unsigned m = 0;
#pragma unroll
            for (int i = 0; i < kPHaloIters; ++i) {
                v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (g_off[i] >= 0) { v[i] = __ldg(reinterpret_cast<const float4*>(d.x + (g_off[i] + cc * 32))); m |= 1u << i; }
            }
            cc_out = cc; vm = m;
        };
        auto store_item = [&](int f, const float4 (&v)[kPHaloIters], int cc, unsigned vm) {
            const int buf = f & 1; const uint32_t ph = (f >> 1) & 1;
            float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
            if (d.in_scale) {
                sc = __ldg(reinterpret_cast<const float4*>(d.in_scale + cc * 32 + chunk * 4));
                sh = __ldg(reinterpret_cast<const float4*>(d.in_shift + cc * 32 + chunk * 4));
            }
            mbar_wait(patch_empty(buf), ph ^ 1u);
            uint8_t* hi_img = sgen + buf * Cfg::PATCH;
            uint8_t* lo_img = hi_img + Cfg::IMG;
#pragma unroll
            for (int i = 0; i < kPHaloIters; ++i) {
                if (p_rc[i] < 0) continue;
                float4 tv = v[i];
                if (d.in_scale && (vm & (1u << i))) {
                    tv.x = fmaf(tv.x, sc.x, sh.x); tv.y = fmaf(tv.y, sc.y, sh.y); tv.z = fmaf(tv.z, sc.z, sh.z); tv.w = fmaf(tv.w, sc.w, sh.w);
                    if (d.in_relu) { tv.x = fmaxf(tv.x, 0.f); tv.y = fmaxf(tv.y, 0.f); tv.z = fmaxf(tv.z, 0.f); tv.w = fmaxf(tv.w, 0.f); }
                }
                Opnd<NPASS>::store(hi_img, lo_img, s_off[i], tv);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(patch_full(buf));
        };
        float4 va[kPHaloIters], vb[kPHaloIters];
        int ca = 0, cb = 0; unsigned ma = 0, mb = 0;
        load_next(va, ca, ma);
        for (int f = 0; f < total; f += 2) {
            load_next(vb, cb, mb);
            store_item(f, va, ca, ma);
            if (f + 1 < total) {
                load_next(va, ca, ma);
                store_item(f + 1, vb, cb, mb);
            }
        }
    } else if (warp == MMA_WARP) {
        // converged warp, one elected lane issues; descriptors = templates advanced by adds (see conv_halo_tma.cu)
        using Op = Opnd<NPASS>;
        const uint32_t idesc = Op::idesc(BN);
        const uint32_t idesc2 = Op::idesc(2 * BN);
        const uint64_t a_tmpl = Op::desc(0, kPPitch * Op::ROW), b_tmpl = Op::desc(0);
        const bool leader = elect_one();
        int f = 0, g = 0;                     // flat patch / weight-stage counters
        for (int ti = 0; ti < my_tiles; ++ti) {
            const int abuf = ti & 1; const uint32_t tph = (ti >> 1) & 1;
            mbar_wait(tmem_empty(abuf), tph ^ 1u);
            tc_fence_after();
            int kb = 0;
            for (int cc = 0; cc < nchunk; ++cc, ++f) {
                const int buf = f & 1; const uint32_t pph = (f >> 1) & 1;
                mbar_wait(patch_full(buf), pph);
                const uint64_t a_hi_d = a_tmpl + (uint64_t)((sbase + buf * Cfg::PATCH) >> 4), a_lo_d = a_hi_d + (uint64_t)(Cfg::IMG >> 4);
                for (int tap = 0; tap < 9; ++tap, ++kb, ++g) {
                    const int s = g % NSTB; const uint32_t ph = (g / NSTB) & 1;
                    mbar_wait(b_full(s), ph);
                    tc_fence_after();
                    if (leader) {
                        const int ky = tap / 3, kx = tap - ky * 3;
                        const uint64_t shift = (uint64_t)(((ky * kPPitch + kx) * Op::ROW) >> 4);
                        const uint64_t b_hi_d = b_tmpl + (uint64_t)((b_base + s * Cfg::B_STAGE) >> 4), b_lo_d = b_hi_d + (uint64_t)((BN * 128) >> 4);
                        const uint32_t acc = tmem + (uint32_t)(abuf * Cfg::BUF_COLS + (kb % NACC) * Cfg::ACC_COLS);
                        const uint32_t fresh = (kb < NACC) ? 0u : 1u;
#pragma unroll
                        for (int kk = 0; kk < Op::KSTEPS; ++kk) {
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
                                Op::mma(acc, dah, dbh, idesc, (kk ? 1u : fresh));
                            }
                        }
                        mma_commit(b_empty(s));
                    }
                    __syncwarp();
                }
                if (leader) mma_commit(patch_empty(buf));
            }
            if (leader) mma_commit(tmem_full(abuf));
            __syncwarp();
        }
    } else if (warp == LOAD_WARP) {
        if (lane == 0) {
            constexpr uint32_t BYTES = Cfg::B_STAGE;
            const int nkb = nchunk * 9;
            int g = 0;
            for (int ti = 0; ti < my_tiles; ++ti) {
                const int nt = ((int)blockIdx.x + ti * (int)gridDim.x) % p.ntile_n;
                const uint8_t* src = reinterpret_cast<const uint8_t*>(p.wt) + (size_t)nt * nkb * BYTES;
                for (int kb = 0; kb < nkb; ++kb, ++g) {
                    const int s = g % NSTB; const uint32_t ph = (g / NSTB) & 1;
                    mbar_wait(b_empty(s), ph ^ 1u);
                    mbar_expect_tx(b_full(s), BYTES);
                    bulk_g2s(b_base + s * Cfg::B_STAGE, src + (size_t)kb * BYTES, BYTES, b_full(s));
                }
            }
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
        for (int ti = 0; ti < my_tiles; ++ti) {
            const int abuf = ti & 1; const uint32_t tph = (ti >> 1) & 1;
            int b, y0, x0, n0; tile_coords(ti, b, y0, x0, n0);
            const int oy = y0 + (row >> 3), ox = x0 + (row & 7);
            const size_t m = (size_t)(b * d.Hout + oy) * d.Wout + ox;
            float* yp = d.y + m * d.y_ld;
            const float rs = d.row_scale ? (d.row_scale[m] + d.row_scale_add) : 1.f;
            if (ti > 0 && d.stat_sum) asm volatile("bar.sync 1, 256;" ::: "memory");      // red[] of the previous tile consumed
            mbar_wait(tmem_full(abuf), tph);
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
                for (int i = etid; i < BN; i += kPHaloEpilogue) {
                    if (n0 + i < d.Cout) {
                        float s1 = 0.f, s2 = 0.f;
#pragma unroll
                        for (int w = 0; w < 4; ++w) { s1 += red[(w * 2 + 0) * BN + i]; s2 += red[(w * 2 + 1) * BN + i]; }
                        atomicAdd(d.stat_sum + n0 + i, (double)s1);
                        atomicAdd(d.stat_sumsq + n0 + i, (double)s2);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(Cfg::TMEM_COLS) : "memory");
    }
}

template <int BN, int NPASS>
static int launch_halo_persist(const HaloPP& p0, cudaStream_t st) {
    using Cfg = HaloPCfg<BN, NPASS>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_halo_persist_kernel<BN, NPASS>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
        if (e != cudaSuccess) { set_error("conv_halo: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); return SAUNET_ERR_CUDA; }
        attr_set = true;
    }
    HaloPP p = p0;
    p.ntile_n = cdiv(p.d.Cout, BN);
    p.ntiles = p.d.B * p.tiles_y * p.tiles_x * p.ntile_n;
    const int grid = p.ntiles < kNumSMs ? p.ntiles : kNumSMs;
    conv_halo_persist_kernel<BN, NPASS><<<grid, kPHaloThreads, Cfg::SMEM, st>>>(p);
    SAUNET_CHECK_LAUNCH("conv_halo_persist_kernel");
    return SAUNET_OK;
}

// entry for conv_fwd_halo (conv_halo.cu): BN = 128 tiles only -- with one producer group per SM the narrow tiles (whose
// MMAs are short) are bound by the latency of the patch gathers and stay on the two-CTAs-per-SM kernel
int conv_fwd_halo_persist(const saunet_conv_desc* d, cudaStream_t st) {
    HaloPP p; p.d = *d;
    p.tiles_x = d->Win / 8; p.tiles_y = d->Hin / 16; p.nchunk = d->Cin / 32; p.wt = d->w_tc;
    if (d->tc_passes == kBF16) return launch_halo_persist<128, kBF16>(p, st);
    return d->tc_passes != 1 ? launch_halo_persist<128, 3>(p, st) : launch_halo_persist<128, 1>(p, st);
}

}  // namespace saunet
