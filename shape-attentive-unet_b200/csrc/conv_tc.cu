// conv_tc.cu -- implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05.mma, TMEM accumulators).
//
// GEMM view (same descriptor as conv_simt.cu): rows = 128 output pixels per tile, N = Cout tile (16..256),
// K = KH*KW*Cin in blocks of 32 fp32 (= one 128-byte swizzle row).
//
// PERSISTENT kernel: one CTA per SM walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...  All roles run flat loops
// over (tile, k-block), so the pipelines never drain at a tile boundary and the epilogue of tile i overlaps the main
// loop of tile i+1 through two TMEM accumulator buffers (the DenseNet 1x1 layers have only 2..31 k-blocks per tile:
// a one-tile-per-CTA kernel spends half its life in prologue + epilogue; ncu: 14 % tensor-active).
//
// Warp roles (832 threads):
//   warps 0-15   A producers: gather the im2col rows from NHWC global memory (128-bit loads), apply the fused
//                BatchNorm+ReLU prologue, split every value into tf32 hi + lo, store both tiles into shared
//                memory in the UMMA K-major SWIZZLE_128B layout, fence.proxy.async, arrive on full_a[stage].
//                A 3-deep register ring keeps two k-blocks of gathers in flight, across tile boundaries too.
//   warps 16-23  epilogue: tcgen05.ld -> bias / BN-statistics / gate row-scale / activation -> global, then
//                release the accumulator buffer (tmem_empty).
//   warp 24      allocates TMEM; one elected lane issues tcgen05.mma.kind::tf32 (M=128, N=BN, K=8):
//                3xTF32 = a_lo*b_hi + a_hi*b_lo + a_hi*b_hi with fp32 accumulation in TMEM (error ~2^-21, the
//                accuracy class the 1e-4 parity gate needs), or one pass (plain TF32) on request.
//   warp 25      one lane streams the pre-tiled, pre-swizzled, pre-split weight tiles with cp.async.bulk (TMA
//                engine, 1-D bulk copy) completing on full_b[stage].
// Pipeline: NSTAGE-deep smem ring (mbarrier full/empty per stage, tcgen05.commit releases a stage) + 2 accumulator
// buffers (tmem_full / tmem_empty).
#include "tc_common.cuh"
#include <stdlib.h>

extern "C" int saunet_tc_chunk_major(int taps, int Cin);

namespace saunet {

struct TcP {
    saunet_conv_desc d;
    int M, K, HgWg, nkb;
    int ntile_n, ntiles;
    int chunk_major;      // K ordered (32-channel chunk, tap, channel) instead of (tap, channel): see pack_tc_kernel
    const float* wt;      // tiled weights: [n_tile][k_block][pass][BN][32] (swizzled image)
    long long* prof;      // optional per-CTA cycle counters (tools/bench_conv.py, SAUNET_TC_PROF)
};

template <int BN, int NPASS>
struct TcCfg {
    using Op = Opnd<NPASS>;
    static constexpr int A_BYTES = 128 * Op::ROW;
    static constexpr int B_BYTES = BN * Op::ROW;
    static constexpr int NOP = Op::NOP;                              // hi (+ lo) images per operand (bf16: one image of 64-byte rows)
    static constexpr int STAGE = NOP * (A_BYTES + B_BYTES);
    static constexpr int RED_BYTES = 8 * BN * 4;                     // [4 lane quarters][sum, sumsq][BN] floats
    // output staging tile for the bulk-copy epilogue: 128 rows of BN floats at a pitch of BN*4 + 16 bytes (the 16 bytes
    // of padding make the row-per-thread 128-bit stores conflict-free)
    static constexpr int OUT_PITCH = BN * 4 + 16;
    static constexpr int OUT_BYTES = 128 * OUT_PITCH;
    static constexpr int NSTAGE_RAW = (227 * 1024 - 1024 - 256 - RED_BYTES - OUT_BYTES) / STAGE;
    static constexpr int NSTAGE = NSTAGE_RAW > 4 ? 4 : NSTAGE_RAW;
    static_assert(NSTAGE >= 2, "shared memory: two pipeline stages must fit");
    static constexpr int SMEM = NSTAGE * STAGE + OUT_BYTES + RED_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
    static constexpr int ACC_COLS = BN < 32 ? 32 : BN;
    static constexpr int NBUF = 2;                                   // accumulator buffers (epilogue overlap)
    // The tensor core rounds its fp32 accumulator toward zero on every MMA, a bias that grows linearly with K
    // (measured 4.5e-9*K normalised).  Round-robin the k-blocks over NACC TMEM accumulators and add them in the
    // epilogue with round-to-nearest FADDs: the bias drops by ~NACC.
    // (bf16 operands: one accumulator -- the round-robin over several only serves the fp32-class error budget of 3xTF32, and
    //  every extra accumulator is one more TMEM load per 16 columns in an epilogue that bounds the short-K tiles)
    static constexpr int NACC = Op::BF ? 1 : ((512 / (NBUF * ACC_COLS)) > 4 ? 4 : (512 / (NBUF * ACC_COLS)));
    static constexpr int BUF_COLS = NACC * ACC_COLS;
    static constexpr int TMEM_COLS = NBUF * BUF_COLS <= 256 ? 256 : 512;
};

constexpr int kProducers = 512;      // 16 producer warps (warps 0-15): 2 rows x one 16-byte chunk per thread and k-block
constexpr int kEpilogue = 256;       // 8 epilogue warps (warps 16-23)
constexpr int kTcThreads = kProducers + kEpilogue + 64;     // + MMA warp (24) + weight loader warp (25)
constexpr int kRowsPerThread = 128 * 8 / kProducers;        // 2

#define TC_PROF_T0() long long pt0__ = p.prof ? clock64() : 0
#define TC_PROF_ADD(var) do { if (p.prof) { long long t1__ = clock64(); (var) += t1__ - pt0__; pt0__ = t1__; } } while (0)

// MODE 0: generic K order (tap-major, any Cin % 4 == 0; per-chunk tap via integer division)
// MODE 1: pointwise 1x1 / stride 1 / no padding -- no taps, no spatial bounds (the DenseNet bottleneck convs and their dgrads)
// MODE 2: chunk-major K order (taps > 1, Cin % 32 == 0): (chunk, tap) advance incrementally, no divisions
template <int BN, int NPASS, int MODE>
__global__ void __launch_bounds__(kTcThreads, 1) conv_tc_kernel(const __grid_constant__ TcP p) {
    using Cfg = TcCfg<BN, NPASS>;
    constexpr int NSTAGE = Cfg::NSTAGE;
    constexpr int NACC = Cfg::NACC;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
    uint8_t* ostage = sgen + NSTAGE * Cfg::STAGE;
    float* red = reinterpret_cast<float*>(ostage + Cfg::OUT_BYTES);
    const uint32_t bars = sbase + NSTAGE * Cfg::STAGE + Cfg::OUT_BYTES + Cfg::RED_BYTES;
    auto full_a = [&](int s) { return bars + 8u * s; };
    auto full_b = [&](int s) { return bars + 8u * (NSTAGE + s); };
    auto empty = [&](int s) { return bars + 8u * (2 * NSTAGE + s); };
    auto tmem_full = [&](int b) { return bars + 8u * (3 * NSTAGE + b); };
    auto tmem_empty = [&](int b) { return bars + 8u * (3 * NSTAGE + 2 + b); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sgen + NSTAGE * Cfg::STAGE + Cfg::OUT_BYTES + Cfg::RED_BYTES + 8 * (3 * NSTAGE + 4));

    const saunet_conv_desc& d = p.d;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int NPW = kProducers / 32, EPI_WARP0 = NPW, MMA_WARP = NPW + kEpilogue / 32, LOAD_WARP = MMA_WARP + 1;
    constexpr int RPT = kRowsPerThread, RSTEP = 128 / RPT;
    const int my_tiles = (p.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;    // >= 1 (grid <= ntiles)
    const int nkb = p.nkb;

    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(full_a(s), kProducers / 32); mbar_init(full_b(s), 1); mbar_init(empty(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tmem_full(b), 1); mbar_init(tmem_empty(b), kEpilogue / 32); }
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
    const long long kt0 = p.prof ? clock64() : 0;

    if (warp < NPW) {
        // ================= A producers =================
        const int chunk = tid & 7, rbase = tid >> 3;          // RPT rows per thread: rbase + RSTEP*i
        const int taps = d.KH * d.KW;
        const int total = my_tiles * nkb;
        long long pw_empty = 0, pw_store = 0, pw_load = 0;
        // load cursor: (tile index in this CTA's list, k-block) of the next gather + that tile's per-row constants:
        // element offset of (b, iy0, ix0) and the tap-(0,0) coordinates; rows past M fail every bounds test.
        // 32-bit element offsets (host guarantees the tensor fits).  The loop runs flat over (tile, k-block).
        int l_ti = -1, l_kb = 0, l_next_ti = 0;
        int l_tap = 0, l_ky = 0, l_kx = 0, l_cc = 0;          // MODE 2 cursor: tap (ky, kx) within the 32-channel chunk l_cc
        int r_iy0[RPT], r_ix0[RPT], r_base[RPT];
        uint32_t s_off[RPT];                                  // swizzled byte offset of this thread's chunk in each of its rows
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
            const int r = rbase + RSTEP * i;
            s_off[i] = Opnd<NPASS>::off(r, chunk);
        }
        auto set_tile = [&](int ti) {
            const int tile = (int)blockIdx.x + ti * (int)gridDim.x;
            const int m0 = (tile / p.ntile_n) * 128;
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
                const int m = m0 + rbase + RSTEP * i;
                if (m < p.M) {
                    if (MODE == 1) { r_iy0[i] = 0; r_ix0[i] = 0; r_base[i] = m * d.x_ld; }
                    else {
                        int b = m / p.HgWg; int r = m - b * p.HgWg; int gi = r / d.Wg; int gj = r - gi * d.Wg;
                        r_iy0[i] = gi * d.sy + d.offy; r_ix0[i] = gj * d.sx + d.offx;
                        r_base[i] = ((b * d.Hin + r_iy0[i]) * d.Win + r_ix0[i]) * d.x_ld;
                    }
                } else { r_iy0[i] = -(1 << 28); r_ix0[i] = -(1 << 28); r_base[i] = -1; }
            }
            l_ti = ti;
        };
        // gather of the next k-block into registers; `tag` = channel of the thread's chunk (0xFFFFF: K padding) in the low
        // bits and one validity bit per row above them (padding / out-of-image rows stay exactly zero after the prologue)
        auto load_next = [&](float4 (&v)[RPT], int& tag) {
            if (l_next_ti >= my_tiles) return;
            if (l_next_ti != l_ti) set_tile(l_next_ti);
            const int kb = l_kb;
            int c, ky = 0, kx = 0; bool kval = true;
            if (MODE == 1) { c = kb * 32 + chunk * 4; kval = c < d.Cin; }
            else if (MODE == 2) { c = l_cc * 32 + chunk * 4; ky = l_ky; kx = l_kx; }
            else { const int k = kb * 32 + chunk * 4; kval = k < p.K; const int tap = kval ? k / d.Cin : 0; c = k - tap * d.Cin; ky = tap / d.KW; kx = tap - ky * d.KW; }
            if (++l_kb == nkb) { l_kb = 0; ++l_next_ti; l_tap = 0; l_ky = 0; l_kx = 0; l_cc = 0; }
            else if (MODE == 2) {
                if (++l_tap == taps) { l_tap = 0; l_ky = 0; l_kx = 0; ++l_cc; }
                else if (++l_kx == d.KW) { l_kx = 0; ++l_ky; }
            }
            const int tapoff = (MODE == 1) ? c : (ky * d.Win + kx) * d.x_ld + c;
            int t = kval ? c : 0xFFFFF;
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
                bool ok;
                if (MODE == 1) ok = kval && r_base[i] >= 0;
                else ok = kval && (unsigned)(r_iy0[i] + ky) < (unsigned)d.Hin && (unsigned)(r_ix0[i] + kx) < (unsigned)d.Win;
                v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ok) { v[i] = __ldg(reinterpret_cast<const float4*>(d.x + (r_base[i] + tapoff))); t |= (1 << (20 + i)); }
            }
            tag = t;
        };
        auto store_item = [&](int f, const float4 (&v)[RPT], int tag) {
            const int s = f % NSTAGE; const uint32_t ph = (f / NSTAGE) & 1;
            const int cch = tag & 0xFFFFF;
            float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
            if (d.in_scale && cch != 0xFFFFF) {
                sc = __ldg(reinterpret_cast<const float4*>(d.in_scale + cch));
                sh = __ldg(reinterpret_cast<const float4*>(d.in_shift + cch));
            }
            TC_PROF_T0();
            mbar_wait(empty(s), ph ^ 1u);
            TC_PROF_ADD(pw_empty);
            uint8_t* a_hi = sgen + s * Cfg::STAGE;
            uint8_t* a_lo = a_hi + Cfg::A_BYTES;
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
                float4 t = v[i];
                if (d.in_scale && (tag & (1 << (20 + i)))) {
                    t.x = fmaf(t.x, sc.x, sh.x); t.y = fmaf(t.y, sc.y, sh.y); t.z = fmaf(t.z, sc.z, sh.z); t.w = fmaf(t.w, sc.w, sh.w);
                    if (d.in_relu) { t.x = fmaxf(t.x, 0.f); t.y = fmaxf(t.y, 0.f); t.z = fmaxf(t.z, 0.f); t.w = fmaxf(t.w, 0.f); }
                }
                Opnd<NPASS>::store(a_hi, a_lo, s_off[i], t);
            }
            TC_PROF_ADD(pw_load);          // (first use of the gathered registers: global-load latency lands here)
            // every writer fences its own generic-proxy stores towards the async proxy, the warp converges, and ONE
            // lane arrives (per-thread arrivals on one mbarrier serialise in shared memory)
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(full_a(s));
            TC_PROF_ADD(pw_store);
        };
        // software pipeline: a 3-deep register ring keeps two k-blocks of gathers in flight while one is stored
        float4 v0[RPT], v1[RPT], v2[RPT];
        int cc0 = 0, cc1 = 0, cc2 = 0;
        load_next(v0, cc0);
        load_next(v1, cc1);
        for (int f = 0; f < total; f += 3) {
            load_next(v2, cc2);
            store_item(f, v0, cc0);
            if (f + 1 < total) { load_next(v0, cc0); store_item(f + 1, v1, cc1); }
            if (f + 2 < total) { load_next(v1, cc1); store_item(f + 2, v2, cc2); }
        }
        if (p.prof && tid == 0) {
            long long* o = p.prof + blockIdx.x * 16;
            o[0] = clock64() - kt0; o[1] = pw_empty; o[2] = pw_load; o[3] = pw_store;
        }
    } else if (warp == MMA_WARP) {
        // ================= MMA issuer =================
        // converged warp, one elected lane issues; descriptors = templates advanced by adds (see conv_halo_tma.cu: with the
        // loop inside `if (lane == 0)` every tcgen05.mma cost the issuing thread 70-170 clocks)
        {
            // instruction descriptor: D=f32, A=B=tf32 (bf16), both K-major, N=BN, M=128
            using Op = Opnd<NPASS>;
            const uint32_t idesc = Op::idesc(BN);
            const uint64_t tmpl = Op::desc(0);
            const bool leader = elect_one();
            long long mw_te = 0, mw_fa = 0, mw_fb = 0, mw_issue = 0;
            int f = 0;
            for (int ti = 0; ti < my_tiles; ++ti) {
                const int buf = ti & 1; const uint32_t tph = (ti >> 1) & 1;
                TC_PROF_T0();
                mbar_wait(tmem_empty(buf), tph ^ 1u);          // the epilogue has drained this buffer (first use passes)
                TC_PROF_ADD(mw_te);
                tc_fence_after();
                for (int kb = 0; kb < nkb; ++kb, ++f) {
                    const int s = f % NSTAGE; const uint32_t ph = (f / NSTAGE) & 1;
                    mbar_wait(full_a(s), ph);
                    TC_PROF_ADD(mw_fa);
                    mbar_wait(full_b(s), ph);
                    TC_PROF_ADD(mw_fb);
                    tc_fence_after();
                    if (leader) {
                        const uint64_t a_hi = tmpl + (uint64_t)((sbase + s * Cfg::STAGE) >> 4);
                        const uint64_t a_lo = a_hi + (uint64_t)(Cfg::A_BYTES >> 4);
                        const uint64_t b_hi = a_hi + (uint64_t)((Cfg::NOP * Cfg::A_BYTES) >> 4);
                        const uint64_t b_lo = b_hi + (uint64_t)(Cfg::B_BYTES >> 4);
                        const uint32_t acc = tmem + (uint32_t)(buf * Cfg::BUF_COLS + (kb % NACC) * Cfg::ACC_COLS);
                        const uint32_t fresh = (kb < NACC) ? 0u : 1u;          // first k-block of each accumulator overwrites
#pragma unroll
                        for (int kk = 0; kk < Op::KSTEPS; ++kk) {
                            const uint64_t dah = a_hi + (uint64_t)(kk * 2), dbh = b_hi + (uint64_t)(kk * 2);
                            if (NPASS == 3) {
                                const uint64_t dal = a_lo + (uint64_t)(kk * 2), dbl = b_lo + (uint64_t)(kk * 2);
                                mma_tf32(acc, dal, dbh, idesc, (kk ? 1u : fresh));       // small terms first
                                mma_tf32(acc, dah, dbl, idesc, 1u);
                                mma_tf32(acc, dah, dbh, idesc, 1u);
                            } else {
                                Op::mma(acc, dah, dbh, idesc, (kk ? 1u : fresh));
                            }
                        }
                        mma_commit(empty(s));
                    }
                    __syncwarp();
                    TC_PROF_ADD(mw_issue);
                }
                if (leader) mma_commit(tmem_full(buf));
                __syncwarp();
            }
            if (p.prof && leader) {
                long long* o = p.prof + blockIdx.x * 16;
                o[4] = mw_te; o[5] = mw_fa; o[6] = mw_fb; o[7] = mw_issue;
            }
        }
        __syncwarp();
    } else if (warp == LOAD_WARP) {
        // ================= weight-tile loader =================
        if (lane == 0) {
            constexpr uint32_t BYTES = Cfg::NOP * Cfg::B_BYTES;
            int f = 0;
            for (int ti = 0; ti < my_tiles; ++ti) {
                const int tile = (int)blockIdx.x + ti * (int)gridDim.x;
                const int nt = tile % p.ntile_n;
                const uint8_t* src = reinterpret_cast<const uint8_t*>(p.wt) + (size_t)nt * nkb * BYTES;
                for (int kb = 0; kb < nkb; ++kb, ++f) {
                    const int s = f % NSTAGE; const uint32_t ph = (f / NSTAGE) & 1;
                    mbar_wait(empty(s), ph ^ 1u);
                    mbar_expect_tx(full_b(s), BYTES);
                    bulk_g2s(sbase + s * Cfg::STAGE + Cfg::NOP * Cfg::A_BYTES, src + (size_t)kb * BYTES, BYTES, full_b(s));
                }
            }
        }
        __syncwarp();
    } else {
        // ================= epilogue =================
        const int q = warp & 3, half = (warp - EPI_WARP0) >> 2;          // TMEM lane quarter (= warp % 4), column half
        const int etid = tid - EPI_WARP0 * 32;
        const int row = q * 32 + lane;
        // vst: rows are 16-byte aligned -> the tile is staged in shared memory and written (or added, for accumulating
        // data gradients) one whole row per cp.async.bulk: a warp-wide "row per thread" store touches 32 different
        // rows and halves the SM->L2 write rate (the epilogue then bounds every short-K tile)
        const bool vst = (d.Cout % 4 == 0) && (d.y_ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(d.y) & 15u) == 0);
        const int nacc = nkb < NACC ? nkb : NACC;
        long long ew_full = 0, ew_work = 0, ew_ld = 0, ew_alu = 0, ew_bar = 0;
        uint8_t* orow = ostage + (size_t)row * Cfg::OUT_PITCH;
        for (int ti = 0; ti < my_tiles; ++ti) {
            if (ti > 0) {
                // the previous tile's bulk copies have finished reading the staging tile; red[] has been consumed
                if (vst) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
            const int buf = ti & 1; const uint32_t tph = (ti >> 1) & 1;
            const int tile = (int)blockIdx.x + ti * (int)gridDim.x;
            const int m0 = (tile / p.ntile_n) * 128, n0 = (tile % p.ntile_n) * BN;
            const int m = m0 + row;
            const bool mval = m < p.M;
            const bool tile_full = m0 + 128 <= p.M;
            size_t opix = 0;
            if (mval) {
                if (MODE == 1) opix = (size_t)m;
                else {
                    int b = m / p.HgWg; int rr = m - b * p.HgWg; int gi = rr / d.Wg; int gj = rr - gi * d.Wg;
                    opix = (size_t)(b * d.Hout + gi * d.osy + d.oy0) * d.Wout + (gj * d.osx + d.ox0);
                }
            }
            float* yp = d.y + opix * d.y_ld;
            const float rs = (d.row_scale && mval) ? (d.row_scale[m] + d.row_scale_add) : 1.f;
            TC_PROF_T0();
            mbar_wait(tmem_full(buf), tph);
            TC_PROF_ADD(ew_full);
            tc_fence_after();
            const uint32_t tb = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * Cfg::BUF_COLS);
            for (int c0 = half * 16; c0 < BN; c0 += 32) {
                if (n0 + c0 >= d.Cout) break;                     // warp-uniform
                float v[16];
                long long pq0 = p.prof ? clock64() : 0;
                tmem_ld16(tb + (uint32_t)c0, v);
                for (int a = 1; a < nacc; ++a) {
                    float u[16];
                    tmem_ld16(tb + (uint32_t)(a * Cfg::ACC_COLS + c0), u);
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] += u[j];
                }
                if (p.prof) { long long t = clock64(); ew_ld += t - pq0; pq0 = t; }
                float o[16];
                {
                    const int nvalid = d.Cout - (n0 + c0);
                    epi_chunk(v, o, d.bias ? d.bias + n0 + c0 : nullptr, nvalid, tile_full, mval, d.row_scale != nullptr, rs, d.act);
                }
                if (vst) {
#pragma unroll
                    for (int qq = 0; qq < 4; ++qq)
                        *reinterpret_cast<float4*>(orow + (c0 + 4 * qq) * 4) = make_float4(o[4 * qq], o[4 * qq + 1], o[4 * qq + 2], o[4 * qq + 3]);
                } else if (mval) {
                    {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int n = n0 + c0 + j;
                            if (n < d.Cout) yp[n] = d.accumulate ? yp[n] + o[j] : o[j];
                        }
                    }
                }
                if (p.prof) { long long t = clock64(); ew_alu += t - pq0; pq0 = t; }
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
            // the accumulator buffer is free as soon as every epilogue warp has read it
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty(buf));
            long long pq1 = p.prof ? clock64() : 0;
            if (vst) fence_proxy_async();                  // staging-tile stores -> visible to the bulk-copy engine
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (p.prof) ew_bar += clock64() - pq1;
            if (vst && etid < 128 && mval) {               // thread etid < 128 owns row etid (== row)
                const int ncols = (d.Cout - n0) < BN ? (d.Cout - n0) : BN;
                const uint32_t src = smem_u32(orow);
                float* dst = yp + n0;
                if (d.accumulate)
                    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(ncols * 4) : "memory");
                else
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(ncols * 4) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            if (d.stat_sum) {
                for (int i = etid; i < BN; i += kEpilogue) {
                    if (n0 + i < d.Cout) {
                        float s1 = 0.f, s2 = 0.f;
#pragma unroll
                        for (int w = 0; w < 4; ++w) { s1 += red[(w * 2 + 0) * BN + i]; s2 += red[(w * 2 + 1) * BN + i]; }
                        atomicAdd(d.stat_sum + n0 + i, (double)s1);
                        atomicAdd(d.stat_sumsq + n0 + i, (double)s2);
                    }
                }
            }
            TC_PROF_ADD(ew_work);
        }
        if (vst) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        if (p.prof && etid == 0) {
            long long* o = p.prof + blockIdx.x * 16;
            o[8] = ew_full; o[9] = ew_work; o[10] = ew_ld; o[11] = ew_alu; o[12] = ew_bar;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(Cfg::TMEM_COLS) : "memory");
    }
}

// ---- weight tiling: [K][N] row-major (the SIMT packing) -> [n_tile][k_block][pass][BN][32] swizzled image ----
// K order of the tiled image: tap-major (k = tap*Cin + c, the SIMT order) or, when taps > 1 and Cin % 32 == 0,
// chunk-major (k-block = cchunk*taps + tap): the 9 taps of a 3x3 conv over one 32-channel chunk become consecutive
// k-blocks, so a CTA re-reads the same ~(tile+halo) x 128 B of activations 9 times in a row -- from L1.
__global__ void pack_tc_kernel(const float* __restrict__ kn, int K, int N, int BN, int npass, int nkb, int taps, int Cin,
                               int chunk_major, float* __restrict__ out) {
    const int nop = npass == 3 ? 2 : 1;
    const long long tile_f = (long long)BN * 32;                    // floats per image
    const int ntile = (N + BN - 1) / BN;
    const long long total = (long long)ntile * nkb * nop * tile_f;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        long long r = idx;
        const int e = (int)(r % 4); r /= 4;           // element within the 16-byte chunk
        const int pc = (int)(r % 8); r /= 8;          // physical chunk within the 128-byte row
        const int row = (int)(r % BN); r /= BN;
        const int op = (int)(r % nop); r /= nop;
        const int kb = (int)(r % nkb); r /= nkb;
        const int nt = (int)r;
        const int lc = pc ^ (row & 7);                // logical chunk
        int k = kb * 32 + lc * 4 + e;
        bool kval = k < K;
        if (chunk_major) {          // (chunk, tap) order; a partial last chunk (Cin % 32 != 0, padded image) is zero-filled
            const int cc = kb / taps, tap = kb - cc * taps, cch = cc * 32 + lc * 4 + e;
            k = tap * Cin + cch; kval = cch < Cin;
        }
        const int n = nt * BN + row;
        float w = (kval && n < N) ? kn[(size_t)k * N + n] : 0.f;
        float hi = tf32_hi(w);
        out[idx] = (op == 0) ? hi : tf32_hi(w - hi);
    }
}

// bf16 image: [n_tile][k_block][BN][32 bf16] -- 64-byte rows in the K-major SWIZZLE_64B pattern (see Opnd<kBF16>)
__global__ void pack_tc_bf16_kernel(const float* __restrict__ kn, int K, int N, int BN, int nkb, int taps, int Cin,
                                    int chunk_major, uint16_t* __restrict__ out) {
    const int ntile = (N + BN - 1) / BN;
    const long long total = (long long)ntile * nkb * BN * 32;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        long long r = idx;
        const int e = (int)(r % 8); r /= 8;           // element within the 16-byte chunk
        const int pc = (int)(r % 4); r /= 4;          // physical chunk within the 64-byte row
        const int row = (int)(r % BN); r /= BN;
        const int kb = (int)(r % nkb); r /= nkb;
        const int nt = (int)r;
        const int lc = pc ^ ((row >> 1) & 3);         // logical chunk
        int k = kb * 32 + lc * 8 + e;
        bool kval = k < K;
        if (chunk_major) {
            const int cc = kb / taps, tap = kb - cc * taps, cch = cc * 32 + lc * 8 + e;
            k = tap * Cin + cch; kval = cch < Cin;
        }
        const int n = nt * BN + row;
        const float w = (kval && n < N) ? kn[(size_t)k * N + n] : 0.f;
        uint32_t b;
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(b) : "f"(0.f), "f"(w));
        out[idx] = (uint16_t)(b & 0xFFFFu);
    }
}

template <int BN, int NPASS, int MODE>
static int launch_tc_mode(const TcP& p, cudaStream_t st) {
    using Cfg = TcCfg<BN, NPASS>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<BN, NPASS, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
        if (e != cudaSuccess) { set_error("conv_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); return SAUNET_ERR_CUDA; }
        attr_set = true;
    }
    TcP q = p;
    q.ntile_n = cdiv(p.d.Cout, BN);
    q.ntiles = cdiv(p.M, 128) * q.ntile_n;
    // debugging aid: device pointer to long long[148][16] (read once per process)
    static long long* const prof_ptr = []() -> long long* {
        const char* pe = getenv("SAUNET_TC_PROF");
        return pe ? reinterpret_cast<long long*>(strtoull(pe, nullptr, 0)) : nullptr;
    }();
    q.prof = prof_ptr;
    const int grid = q.ntiles < kNumSMs ? q.ntiles : kNumSMs;
    conv_tc_kernel<BN, NPASS, MODE><<<grid, kTcThreads, Cfg::SMEM, st>>>(q);
    SAUNET_CHECK_LAUNCH("conv_tc_kernel");
    return SAUNET_OK;
}
template <int BN, int NPASS>
static int launch_tc(const TcP& p, cudaStream_t st) {
    const saunet_conv_desc& d = p.d;
    if (d.KH == 1 && d.KW == 1 && d.sy == 1 && d.sx == 1 && d.offy == 0 && d.offx == 0 && d.Hg == d.Hin && d.Wg == d.Win)
        return launch_tc_mode<BN, NPASS, 1>(p, st);
    if (p.chunk_major) return launch_tc_mode<BN, NPASS, 2>(p, st);
    return launch_tc_mode<BN, NPASS, 0>(p, st);
}

bool conv_tc_eligible(const saunet_conv_desc* d) {
    if (!d->w_tc || d->tc_bn <= 0) return false;
    if (d->Cin % 4 || d->x_ld % 4 || !aligned16(d->x)) return false;
    if (d->in_scale && (!aligned16(d->in_scale) || !aligned16(d->in_shift))) return false;
    if (!aligned16(d->w_tc)) return false;
    if ((long long)d->B * d->Hin * d->Win * d->x_ld >= (1ll << 31)) return false;      // 32-bit element offsets
    return true;
}

int conv_fwd_tc(const saunet_conv_desc* d, cudaStream_t st) {
    TcP p; p.d = *d;
    long long M = (long long)d->B * d->Hg * d->Wg;
    SAUNET_CHECK_ARG(M > 0 && M < (1ll << 31), SAUNET_ERR_BAD_SHAPE, "conv2d_fwd(tc): bad M=%lld", M);
    p.M = (int)M; p.K = d->KH * d->KW * d->Cin; p.HgWg = d->Hg * d->Wg; p.nkb = (p.K + 31) / 32; p.wt = d->w_tc;
    p.chunk_major = saunet_tc_chunk_major(d->KH * d->KW, d->Cin);
    const int np = d->tc_passes;
    switch (d->tc_bn) {
        case 16: return np == kBF16 ? launch_tc<16, kBF16>(p, st) : np != 1 ? launch_tc<16, 3>(p, st) : launch_tc<16, 1>(p, st);
        case 32: return np == kBF16 ? launch_tc<32, kBF16>(p, st) : np != 1 ? launch_tc<32, 3>(p, st) : launch_tc<32, 1>(p, st);
        case 64: return np == kBF16 ? launch_tc<64, kBF16>(p, st) : np != 1 ? launch_tc<64, 3>(p, st) : launch_tc<64, 1>(p, st);
        case 128: return np == kBF16 ? launch_tc<128, kBF16>(p, st) : np != 1 ? launch_tc<128, 3>(p, st) : launch_tc<128, 1>(p, st);
    }
    set_error("conv2d_fwd(tc): unsupported N tile %d", d->tc_bn);
    return SAUNET_ERR_BAD_SHAPE;
}

}  // namespace saunet

using namespace saunet;

extern "C" int saunet_tc_chunk_major(int taps, int Cin) { return (taps > 1 && Cin % 32 == 0) ? 1 : 0; }
extern "C" int saunet_tc_tile_n(int Cout) {
    if (Cout <= 16) return 16;
    if (Cout <= 32) return 32;
    if (Cout <= 64) return 64;
    return 128;        // (a 256-wide tile leaves no shared memory for the output staging tile of the bulk-copy epilogue)
}
// chunk-major PADDED image (conv_halo_tma.cu): k-blocks ordered (32-channel chunk, tap), the last chunk zero-padded
extern "C" long long saunet_tc_packed_floats_cm(int taps, int Cin, int N, int BN, int passes) {
    if (taps <= 0 || Cin <= 0 || N <= 0 || BN <= 0) return 0;
    const long long nkb = (long long)((Cin + 31) / 32) * taps, ntile = (N + BN - 1) / BN;
    return passes == kBF16 ? ntile * nkb * BN * 16 : ntile * nkb * (passes == 3 ? 2 : 1) * BN * 32;
}
extern "C" int saunet_pack_weights_tc_cm(const float* kn, int taps, int Cin, int N, int BN, int passes, float* out, void* stream) {
    SAUNET_CHECK_ARG(kn && out && taps > 0 && Cin > 0 && N > 0, SAUNET_ERR_BAD_SHAPE, "pack_weights_tc_cm: bad args");
    SAUNET_CHECK_ARG(BN == 16 || BN == 32 || BN == 64 || BN == 128, SAUNET_ERR_BAD_SHAPE, "pack_weights_tc_cm: bad N tile %d", BN);
    SAUNET_CHECK_ARG(passes == 1 || passes == 3 || passes == kBF16, SAUNET_ERR_BAD_SHAPE, "pack_weights_tc_cm: passes must be 1, 3 or 16 (bf16)");
    const long long total = saunet_tc_packed_floats_cm(taps, Cin, N, BN, passes) * (passes == kBF16 ? 2 : 1);
    int blocks = (int)((total + 255) / 256); if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    if (passes == kBF16)
        pack_tc_bf16_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(kn, taps * Cin, N, BN, ((Cin + 31) / 32) * taps, taps, Cin, 1, reinterpret_cast<uint16_t*>(out));
    else
        pack_tc_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(kn, taps * Cin, N, BN, passes, ((Cin + 31) / 32) * taps, taps, Cin, 1, out);
    SAUNET_CHECK_LAUNCH("pack_tc_kernel");
    return SAUNET_OK;
}
extern "C" long long saunet_tc_packed_floats(int K, int N, int BN, int passes) {
    if (K <= 0 || N <= 0 || BN <= 0) return 0;
    const long long nkb = (K + 31) / 32, ntile = (N + BN - 1) / BN;
    return passes == kBF16 ? ntile * nkb * BN * 16 : ntile * nkb * (passes == 3 ? 2 : 1) * BN * 32;      // (bf16: two elements per float)
}
extern "C" int saunet_pack_weights_tc(const float* kn, int taps, int Cin, int N, int BN, int passes, float* out, void* stream) {
    const int K = taps * Cin;
    SAUNET_CHECK_ARG(kn && out && taps > 0 && Cin > 0 && N > 0, SAUNET_ERR_BAD_SHAPE, "pack_weights_tc: bad args");
    SAUNET_CHECK_ARG(BN == 16 || BN == 32 || BN == 64 || BN == 128, SAUNET_ERR_BAD_SHAPE, "pack_weights_tc: bad N tile %d", BN);
    SAUNET_CHECK_ARG(passes == 1 || passes == 3 || passes == kBF16, SAUNET_ERR_BAD_SHAPE, "pack_weights_tc: passes must be 1, 3 or 16 (bf16)");
    const long long total = saunet_tc_packed_floats(K, N, BN, passes) * (passes == kBF16 ? 2 : 1);
    int blocks = (int)((total + 255) / 256); if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    if (passes == kBF16)
        pack_tc_bf16_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(kn, K, N, BN, (K + 31) / 32, taps, Cin,
                                                                      saunet_tc_chunk_major(taps, Cin), reinterpret_cast<uint16_t*>(out));
    else
        pack_tc_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(kn, K, N, BN, passes, (K + 31) / 32, taps, Cin,
                                                                 saunet_tc_chunk_major(taps, Cin), out);
    SAUNET_CHECK_LAUNCH("pack_tc_kernel");
    return SAUNET_OK;
}
