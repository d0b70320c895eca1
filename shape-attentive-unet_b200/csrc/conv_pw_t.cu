// conv_pw_t.cu -- pointwise (1x1 / stride 1) convolution with the OUTPUT CHANNELS on the TMEM lanes and the PIXELS as
// the MMA N dimension:
//
//     Y^T[co (128 lanes)][pixel (256 columns)] += W[co][k] * X[pixel][k]^T            (3xTF32 or single-pass TF32)
//
// A = the pre-tiled weight image ([co_tile][k_block][hi,lo][128][32], the very image conv_tc.cu streams as its B
// operand for 128-wide tiles), B = a 256-pixel activation tile written by the producers (K-major rows, fused BN+ReLU
// prologue, tf32 hi/lo split).  Against conv_tc.cu (pixels on the lanes, N = 128 channels):
//   * every MMA is M=128 x N=256 x K=8: 12 KB of operand reads per 132 math clocks instead of 8 KB per 66 -- 25 % less
//     shared-memory operand traffic per FLOP, half the MMA instructions and barrier round trips per pixel (the 1x1
//     layers are bound by shared-memory bandwidth: MMA operand reads + producer stores + weight copies);
//   * the epilogue needs no transposition: lane = channel, so for every pixel a warp stores 32 consecutive floats (one
//     128-byte line), the bias is one register, and the BatchNorm statistics are plain per-thread sums over the
//     thread's pixel columns (no shuffles, no shared-memory reduction).
// Prototype measured on B200 (tools/exp/conv1x1_t_proto.cu, M 262144, K 224, Cout 128, no prologue): 0.083 ms =
// 181 TFLOP/s against 0.123 ms for conv_tc.cu.  Persistent, warp-specialised like conv_tc.cu: 16 producer warps, 8
// epilogue warps on the second TMEM accumulator buffer, MMA warp, weight-loader warp.
//
// In the batch-16 step it takes 83 of the 174 launches of conv_tc.cu and runs them at 104 TFLOP/s (conv_tc.cu: 68):
// 54.3 -> 53.4 ms/step.  Used when the problem has >= 74 tiles of 256 pixels x 128 channels and the weights are tiled
// 128 wide; SAUNET_CONV1X1_T=0 switches it off.
#include "tc_common.cuh"
#include <stdlib.h>

namespace saunet {

struct PwtP {
    saunet_conv_desc d;
    int M, nkb, ntile_co, ntiles;      // tile id = pixel tile * ntile_co + co tile
    const float* wt;
};

constexpr int kTNP = 256;                                   // pixels per tile (MMA N)
constexpr int kTProd = 512, kTEpi = 256, kTThreads = kTProd + kTEpi + 64;

template <int NPASS>
struct PwtCfg {
    using Op = Opnd<NPASS>;
    static constexpr int NOP = Op::NOP;
    static constexpr int X_IMG = kTNP * Op::ROW, W_IMG = 128 * Op::ROW;
    static constexpr int STAGE = NOP * (X_IMG + W_IMG);
    static constexpr int NSTAGE = (200 * 1024) / STAGE > 4 ? 4 : (200 * 1024) / STAGE;       // 2 (3xTF32) or 4
    static constexpr int STG = 16 * 128 * 4;                   // fused BN-backward epilogue: 16 pixels x 128 channels staged per chunk
    static constexpr int SMEM = NSTAGE * STAGE + 1024 + 256 + 4 * STG;      // (2 epilogue halves x 2 buffers)
};

template <int NPASS>
__global__ void __launch_bounds__(kTThreads, 1) conv_pw_t_kernel(const __grid_constant__ PwtP p) {
    using Cfg = PwtCfg<NPASS>;
    constexpr int NSTAGE = Cfg::NSTAGE;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
    const uint32_t bars = sbase + NSTAGE * Cfg::STAGE;
    auto full_x = [&](int s) { return bars + 8u * s; };
    auto full_w = [&](int s) { return bars + 8u * (NSTAGE + s); };
    auto empty = [&](int s) { return bars + 8u * (2 * NSTAGE + s); };
    auto tmem_full = [&](int b) { return bars + 8u * (3 * NSTAGE + b); };
    auto tmem_empty = [&](int b) { return bars + 8u * (3 * NSTAGE + 2 + b); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sgen + NSTAGE * Cfg::STAGE + 8 * (3 * NSTAGE + 4));
    float* stg_base = reinterpret_cast<float*>(sgen + NSTAGE * Cfg::STAGE + 256);
    const saunet_conv_desc& d = p.d;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int NPW = kTProd / 32, EPI0 = NPW, MMA_WARP = NPW + kTEpi / 32, LOAD_WARP = MMA_WARP + 1;
    const int my_tiles = (p.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int nkb = p.nkb;
    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(full_x(s), NPW); mbar_init(full_w(s), 1); mbar_init(empty(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tmem_full(b), 1); mbar_init(tmem_empty(b), kTEpi / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp < NPW) {
        // ================= producers: 256 pixel rows x 8 chunks per k-block, 4 items per thread (rows rbase + 64 i) =================
        const int chunk = tid & 7, rbase = tid >> 3;
        uint32_t s_off[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { const int r = rbase + 64 * i; s_off[i] = Opnd<NPASS>::off(r, chunk); }
        const int total = my_tiles * nkb;
        int l_ti = 0, l_kb = 0;
        // `tag`: channel of the thread's chunk (0xFFFFF: K padding) | one validity bit per row (pixels past M stay exactly zero)
        auto load_next = [&](float4 (&v)[4], int& tag) {
            if (l_ti >= my_tiles) return;
            const int tile = (int)blockIdx.x + l_ti * (int)gridDim.x;
            const int m0 = (tile / p.ntile_co) * kTNP + rbase;
            const int c = l_kb * 32 + chunk * 4;
            const bool kval = c < d.Cin;
            int t = kval ? c : 0xFFFFF;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int m = m0 + 64 * i;
                v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (kval && m < p.M) { v[i] = __ldg(reinterpret_cast<const float4*>(d.x + (size_t)m * d.x_ld + c)); t |= 1 << (20 + i); }
            }
            tag = t;
            if (++l_kb == nkb) { l_kb = 0; ++l_ti; }
        };
        auto store_item = [&](int f, const float4 (&v)[4], int tag) {
            const int s = f % NSTAGE; const uint32_t ph = (f / NSTAGE) & 1;
            const int cch = tag & 0xFFFFF;
            float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
            if (d.in_scale && cch != 0xFFFFF) {
                sc = __ldg(reinterpret_cast<const float4*>(d.in_scale + cch));
                sh = __ldg(reinterpret_cast<const float4*>(d.in_shift + cch));
            }
            mbar_wait(empty(s), ph ^ 1u);
            uint8_t* x_hi = sgen + s * Cfg::STAGE;
            uint8_t* x_lo = x_hi + Cfg::X_IMG;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float4 t = v[i];
                if (d.in_scale && (tag & (1 << (20 + i)))) {
                    t.x = fmaf(t.x, sc.x, sh.x); t.y = fmaf(t.y, sc.y, sh.y); t.z = fmaf(t.z, sc.z, sh.z); t.w = fmaf(t.w, sc.w, sh.w);
                    if (d.in_relu) { t.x = fmaxf(t.x, 0.f); t.y = fmaxf(t.y, 0.f); t.z = fmaxf(t.z, 0.f); t.w = fmaxf(t.w, 0.f); }
                }
                Opnd<NPASS>::store(x_hi, x_lo, s_off[i], t);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(full_x(s));
        };
        float4 va[4], vb[4];
        int ta = 0, tb = 0;
        load_next(va, ta);
        for (int f = 0; f < total; f += 2) {
            load_next(vb, tb);
            store_item(f, va, ta);
            if (f + 1 < total) { load_next(va, ta); store_item(f + 1, vb, tb); }
        }
    } else if (warp == MMA_WARP) {
        // converged warp, one elected lane issues; descriptors = templates advanced by adds (see conv_halo_tma.cu)
        {
            // D=f32, A=B=tf32 (bf16), both K-major, N=256 (pixels), M=128 (channels)
            using Op = Opnd<NPASS>;
            const uint32_t idesc = Op::idesc(kTNP);
            const uint64_t tmpl = Op::desc(0);
            const bool leader = elect_one();
            int f = 0;
            for (int ti = 0; ti < my_tiles; ++ti) {
                const int buf = ti & 1; const uint32_t tph = (ti >> 1) & 1;
                mbar_wait(tmem_empty(buf), tph ^ 1u);
                tc_fence_after();
                for (int kb = 0; kb < nkb; ++kb, ++f) {
                    const int s = f % NSTAGE; const uint32_t ph = (f / NSTAGE) & 1;
                    mbar_wait(full_x(s), ph);
                    mbar_wait(full_w(s), ph);
                    tc_fence_after();
                    if (leader) {
                        const uint64_t x_hi = tmpl + (uint64_t)((sbase + s * Cfg::STAGE) >> 4), x_lo = x_hi + (uint64_t)(Cfg::X_IMG >> 4);
                        const uint64_t w_hi = x_hi + (uint64_t)((Cfg::NOP * Cfg::X_IMG) >> 4), w_lo = w_hi + (uint64_t)(Cfg::W_IMG >> 4);
                        const uint32_t acc = tmem + (uint32_t)(buf * kTNP);
#pragma unroll
                        for (int kk = 0; kk < Op::KSTEPS; ++kk) {
                            const uint64_t dwh = w_hi + (uint64_t)(kk * 2), dxh = x_hi + (uint64_t)(kk * 2);
                            if (NPASS == 3) {
                                const uint64_t dwl = w_lo + (uint64_t)(kk * 2), dxl = x_lo + (uint64_t)(kk * 2);
                                mma_tf32(acc, dwl, dxh, idesc, (kb | kk) ? 1u : 0u);
                                mma_tf32(acc, dwh, dxl, idesc, 1u);
                                mma_tf32(acc, dwh, dxh, idesc, 1u);
                            } else {
                                Op::mma(acc, dwh, dxh, idesc, (kb | kk) ? 1u : 0u);
                            }
                        }
                        mma_commit(empty(s));
                    }
                    __syncwarp();
                }
                if (leader) mma_commit(tmem_full(buf));
                __syncwarp();
            }
        }
        __syncwarp();
    } else if (warp == LOAD_WARP) {
        if (lane == 0) {
            constexpr uint32_t BYTES = Cfg::NOP * Cfg::W_IMG;
            int f = 0;
            for (int ti = 0; ti < my_tiles; ++ti) {
                const int ct = ((int)blockIdx.x + ti * (int)gridDim.x) % p.ntile_co;
                const uint8_t* src = reinterpret_cast<const uint8_t*>(p.wt) + (size_t)ct * nkb * BYTES;
                for (int kb = 0; kb < nkb; ++kb, ++f) {
                    const int s = f % NSTAGE; const uint32_t ph = (f / NSTAGE) & 1;
                    mbar_wait(empty(s), ph ^ 1u);
                    mbar_expect_tx(full_w(s), BYTES);
                    bulk_g2s(sbase + s * Cfg::STAGE + Cfg::NOP * Cfg::X_IMG, src + (size_t)kb * BYTES, BYTES, full_w(s));
                }
            }
        }
        __syncwarp();
    } else {
        // ================= epilogue: lane = output channel, columns = pixels =================
        const int q = warp & 3, half = (warp - EPI0) >> 2;
        int bnb_chunks = 0;
        for (int ti = 0; ti < my_tiles; ++ti) {
            const int buf = ti & 1; const uint32_t tph = (ti >> 1) & 1;
            const int tile = (int)blockIdx.x + ti * (int)gridDim.x;
            const int m0 = (tile / p.ntile_co) * kTNP, co = (tile % p.ntile_co) * 128 + q * 32 + lane;
            const bool cov = co < d.Cout;
            const float bj = (d.bias && cov) ? __ldg(d.bias + co) : 0.f;
            const int act = d.act;
            const bool has_rs = d.row_scale != nullptr, accum = d.accumulate != 0;
            float s1 = 0.f, s2 = 0.f;
            // fused BatchNorm(+ReLU) backward epilogue (d.epi_x): this conv is the data gradient that produces
            // da = d loss / d relu(bn(x)); per output channel co (= this thread): g = da * [scale*x + shift > 0],
            // y (+)= scale * g (the data-dependent term of the BatchNorm gradient), stat_sum += sum g,
            // stat_sumsq += sum g * (x - mean).  The two mean terms are applied later from those sums
            // (saunet_bn_fused_finish / saunet_bn_fixup): `da` never goes to HBM and x is read once.
            const bool bnb = d.epi_x != nullptr;
            float e_sc = 0.f, e_sh = 0.f, e_mu = 0.f;
            if (bnb && cov) { e_sc = __ldg(d.epi_scale + co); e_sh = __ldg(d.epi_shift + co); e_mu = __ldg(d.epi_mean + co); }
            mbar_wait(tmem_full(buf), tph);
            tc_fence_after();
            const uint32_t tb = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * kTNP);
            for (int c0 = half * 128; c0 < half * 128 + 128; c0 += 16) {
                const int mc = m0 + c0;
                if (mc >= p.M) break;                           // warp-uniform
                float v[16];
                tmem_ld16(tb + (uint32_t)c0, v);
                const bool full = mc + 16 <= p.M;               // warp-uniform
                if (bnb) {
                    // Output staged as [16 pixels][128 channels] per (half, buffer) and sent as one bulk reduce-add (or plain
                    // bulk store for a first writer) per pixel row, so the epilogue warps only LOAD x.  Measured on B200
                    // (M 262144, K 128, N 224): 0.28 ms; with a register read-modify-write of y (32 loads in flight per
                    // thread, L2-prefetched) 0.40 ms; the same GEMM writing plain da: 0.115 ms + 0.25 ms of reduce / apply passes.
                    float* stg = stg_base + (half * 2 + (bnb_chunks++ & 1)) * (Cfg::STG / 4);      // alternate the two buffers
                    const bool issuer = q == 0 && lane < 16;
                    const float* xp = d.epi_x + (size_t)mc * d.epi_x_ld + co;
                    float xv[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j)                 // 16 independent 128-byte warp requests in flight
                        xv[j] = (cov && (full || mc + j < p.M)) ? __ldg(xp + (size_t)j * d.epi_x_ld) : 0.f;
                    if (lane < 16 && mc + 32 + lane < p.M && co - lane < d.Cout)      // pull the chunk after next into L2 meanwhile (0.38 -> 0.28 ms)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(d.epi_x + (size_t)(mc + 32 + lane) * d.epi_x_ld + (co - lane)) : "memory");
                    if (issuer) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");     // the buffer's previous rows have been read
                    asm volatile("bar.sync %0, 128;" ::"r"(1 + half) : "memory");
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float pre = fmaf(xv[j], e_sc, e_sh);
                        const float g = ((full || mc + j < p.M) && (!d.epi_relu || pre > 0.f)) ? v[j] : 0.f;
                        s1 += g; s2 = fmaf(g, xv[j] - e_mu, s2);
                        stg[j * 128 + q * 32 + lane] = e_sc * g;
                    }
                    fence_proxy_async();
                    asm volatile("bar.sync %0, 128;" ::"r"(1 + half) : "memory");
                    if (issuer && mc + lane < p.M) {
                        const int n0c = (tile % p.ntile_co) * 128;
                        const int ncols = (d.Cout - n0c) < 128 ? (d.Cout - n0c) : 128;
                        float* dst = d.y + (size_t)(mc + lane) * d.y_ld + n0c;
                        const uint32_t src = smem_u32(stg + lane * 128);
                        if (accum)
                            asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(ncols * 4) : "memory");
                        else
                            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(ncols * 4) : "memory");
                    }
                    if (issuer) asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    continue;
                }
                float o[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float pre = (full || mc + j < p.M) ? v[j] + bj : 0.f;      // pixels past M count as nothing
                    s1 += pre; s2 = fmaf(pre, pre, s2);
                    o[j] = pre;
                }
                if (has_rs) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) if (full || mc + j < p.M) o[j] *= __ldg(d.row_scale + mc + j) + d.row_scale_add;
                }
                if (act == SAUNET_ACT_RELU) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) o[j] = fmaxf(o[j], 0.f);
                } else if (act == SAUNET_ACT_SIGMOID) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) o[j] = sigmoid_acc(o[j]);
                }
                if (cov) {
                    float* yp = d.y + (size_t)mc * d.y_ld + co;
                    if (accum) {                                // all 16 loads first (a load-add-store chain per element ran 4x slower)
                        float yv[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) yv[j] = (full || mc + j < p.M) ? yp[(size_t)j * d.y_ld] : 0.f;
#pragma unroll
                        for (int j = 0; j < 16; ++j) o[j] += yv[j];
                    }
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (full || mc + j < p.M) yp[(size_t)j * d.y_ld] = o[j];
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty(buf));
            if (d.stat_sum && cov) {
                atomicAdd(d.stat_sum + co, (double)s1);
                atomicAdd(d.stat_sumsq + co, (double)s2);
            }
        }
        if (d.epi_x) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

template <int NPASS>
static int launch_pwt(const PwtP& p, cudaStream_t st) {
    using Cfg = PwtCfg<NPASS>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_pw_t_kernel<NPASS>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
        if (e != cudaSuccess) { set_error("conv_pw_t: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); return SAUNET_ERR_CUDA; }
        attr_set = true;
    }
    const int grid = p.ntiles < kNumSMs ? p.ntiles : kNumSMs;
    conv_pw_t_kernel<NPASS><<<grid, kTThreads, Cfg::SMEM, st>>>(p);
    SAUNET_CHECK_LAUNCH("conv_pw_t_kernel");
    return SAUNET_OK;
}

bool conv_tc_eligible(const saunet_conv_desc* d);

bool conv_pw_t_eligible(const saunet_conv_desc* d) {
    static const bool off = getenv("SAUNET_CONV1X1_T") != nullptr && getenv("SAUNET_CONV1X1_T")[0] == '0';
    if ((off && !d->epi_x) || !conv_tc_eligible(d) || d->tc_bn != 128) return false;
    if (d->KH != 1 || d->KW != 1 || d->sy != 1 || d->sx != 1 || d->offy != 0 || d->offx != 0) return false;
    if (d->osy != 1 || d->osx != 1 || d->oy0 != 0 || d->ox0 != 0) return false;
    if (d->Hg != d->Hin || d->Wg != d->Win || d->Hout != d->Hin || d->Wout != d->Win) return false;
    const long long M = (long long)d->B * d->Hg * d->Wg;
    if (M >= (1ll << 31)) return false;
    if (d->epi_x) return true;          // the fused BatchNorm-backward epilogue lives in this kernel only
    // 256-pixel tiles: keep at least ~half the SMs busy, otherwise conv_tc.cu's 128-pixel tiles spread better
    return ((M + kTNP - 1) / kTNP) * ((d->Cout + 127) / 128) >= kNumSMs / 2;
}

int conv_fwd_pw_t(const saunet_conv_desc* d, cudaStream_t st) {
    PwtP p; p.d = *d;
    p.M = (int)((long long)d->B * d->Hg * d->Wg);
    p.nkb = (d->Cin + 31) / 32;
    p.ntile_co = cdiv(d->Cout, 128);
    p.ntiles = cdiv(p.M, kTNP) * p.ntile_co;
    p.wt = d->w_tc;
    return d->tc_passes == kBF16 ? launch_pwt<kBF16>(p, st) : d->tc_passes != 1 ? launch_pwt<3>(p, st) : launch_pwt<1>(p, st);
}

}  // namespace saunet
