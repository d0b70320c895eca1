// api.cu -- library-wide state (error string, launch counter, version) and conv dispatch.
#include "common.cuh"
#include <atomic>
#include <string.h>
#include <stdlib.h>

namespace saunet {
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};
void set_error(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
static thread_local const char* g_last_kernel = "";
void note_kernel(const char* name) { g_last_kernel = name; }
const char* last_kernel() { return g_last_kernel; }
int conv_fwd_simt(const saunet_conv_desc* d, cudaStream_t st);
int conv_wgrad_simt(const saunet_wgrad_desc* d, cudaStream_t st);
int conv_fwd_tc(const saunet_conv_desc* d, cudaStream_t st);
bool conv_tc_eligible(const saunet_conv_desc* d);
int conv_fwd_skinny(const saunet_conv_desc* d, cudaStream_t st);
bool conv_skinny_eligible(const saunet_conv_desc* d);
int conv_wgrad_skinny(const saunet_wgrad_desc* d, cudaStream_t st);
bool conv_wgrad_skinny_eligible(const saunet_wgrad_desc* d);
int conv_fwd_halo(const saunet_conv_desc* d, cudaStream_t st);
bool conv_halo_eligible(const saunet_conv_desc* d);
int conv_wgrad_tc(const saunet_wgrad_desc* d, cudaStream_t st);
bool conv_wgrad_tc_eligible(const saunet_wgrad_desc* d);
int conv_wgrad_halo(const saunet_wgrad_desc* d, cudaStream_t st);
bool conv_wgrad_halo_eligible(const saunet_wgrad_desc* d);
int conv_wgrad_pw(const saunet_wgrad_desc* d, cudaStream_t st);
bool conv_wgrad_pw_eligible(const saunet_wgrad_desc* d);
int conv_fwd_pw_t(const saunet_conv_desc* d, cudaStream_t st);
int conv_fwd_halo_tma(const saunet_conv_desc* d, cudaStream_t st);
bool conv_halo_tma_eligible(const saunet_conv_desc* d);
bool conv_pw_t_eligible(const saunet_conv_desc* d);
}  // namespace saunet

using namespace saunet;

extern "C" int saunet_version(void) { return 100; }
extern "C" const char* saunet_last_error(void) { return g_err; }
extern "C" long long saunet_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
extern "C" const char* saunet_last_kernel(void) { return saunet::last_kernel(); }

extern "C" int saunet_conv2d_fwd(const saunet_conv_desc* d, void* stream) {
    SAUNET_CHECK_ARG(d && d->x && d->w && d->y, SAUNET_ERR_BAD_SHAPE, "conv2d_fwd: null pointer");
    SAUNET_CHECK_ARG(d->B > 0 && d->Hin > 0 && d->Win > 0 && d->Cin > 0 && d->Cout > 0 && d->KH > 0 && d->KW > 0 &&
                     d->Hg > 0 && d->Wg > 0 && d->sy > 0 && d->sx > 0 && d->osy > 0 && d->osx > 0,
                     SAUNET_ERR_BAD_SHAPE, "conv2d_fwd: non-positive dimension");
    SAUNET_CHECK_ARG(d->x_ld >= d->Cin && d->y_ld >= d->Cout, SAUNET_ERR_BAD_SHAPE, "conv2d_fwd: ld smaller than channel count");
    SAUNET_CHECK_ARG((d->Hg - 1) * d->osy + d->oy0 < d->Hout && (d->Wg - 1) * d->osx + d->ox0 < d->Wout && d->oy0 >= 0 && d->ox0 >= 0,
                     SAUNET_ERR_BAD_SHAPE, "conv2d_fwd: output grid exceeds output tensor");
    SAUNET_CHECK_ARG((d->in_scale == nullptr) == (d->in_shift == nullptr), SAUNET_ERR_BAD_SHAPE, "conv2d_fwd: in_scale/in_shift mismatch");
    SAUNET_CHECK_ARG((d->stat_sum == nullptr) == (d->stat_sumsq == nullptr), SAUNET_ERR_BAD_SHAPE, "conv2d_fwd: stat_sum/stat_sumsq mismatch");
    SAUNET_CHECK_ARG(d->act >= 0 && d->act <= 2, SAUNET_ERR_BAD_SHAPE, "conv2d_fwd: bad activation %d", d->act);
    if (d->epi_x) {
        SAUNET_CHECK_ARG(d->epi_scale && d->epi_shift && d->epi_mean && d->epi_x_ld >= d->Cout && d->stat_sum && !d->bias && !d->row_scale &&
                         d->act == SAUNET_ACT_NONE && conv_pw_t_eligible(d), SAUNET_ERR_BAD_SHAPE,
                         "conv2d_fwd: the fused BatchNorm-backward epilogue needs a 1x1 / stride-1 tensor-core conv with 128-wide weight "
                         "tiles, statistics targets, and no bias / row scale / activation");
        return conv_fwd_pw_t(d, (cudaStream_t)stream);
    }
    const bool prefer_tc = SAUNET_ENV_FLAG("SAUNET_PREFER_TC") && d->w_tc && d->Cin >= 32 && d->Cout >= 32;
    if (conv_skinny_eligible(d) && !prefer_tc) return conv_fwd_skinny(d, (cudaStream_t)stream);
    if (conv_halo_tma_eligible(d)) return conv_fwd_halo_tma(d, (cudaStream_t)stream);      // 3x3: TMA-fed persistent halo kernel
    SAUNET_CHECK_ARG(!d->tc_cm, SAUNET_ERR_BAD_SHAPE, "conv2d_fwd: chunk-major padded weights (tc_cm) given for a geometry the TMA halo kernel does not take");
    if (conv_halo_eligible(d) && !SAUNET_ENV_FLAG("SAUNET_NO_HALO")) return conv_fwd_halo(d, (cudaStream_t)stream);
    if (conv_pw_t_eligible(d)) return conv_fwd_pw_t(d, (cudaStream_t)stream);      // large 1x1 layers: channels on the TMEM lanes
    if (conv_tc_eligible(d)) return conv_fwd_tc(d, (cudaStream_t)stream);
    return conv_fwd_simt(d, (cudaStream_t)stream);
}

extern "C" int saunet_conv2d_wgrad(const saunet_wgrad_desc* d, void* stream) {
    SAUNET_CHECK_ARG(d && d->p && d->q && d->dw, SAUNET_ERR_BAD_SHAPE, "conv2d_wgrad: null pointer");
    SAUNET_CHECK_ARG(d->B > 0 && d->Hq > 0 && d->Wq > 0 && d->Ca > 0 && d->Cb > 0 && d->KH > 0 && d->KW > 0 && d->Hg > 0 &&
                     d->Wg > 0 && d->sy > 0 && d->sx > 0, SAUNET_ERR_BAD_SHAPE, "conv2d_wgrad: non-positive dimension");
    SAUNET_CHECK_ARG(d->p_ld >= d->Ca && d->q_ld >= d->Cb, SAUNET_ERR_BAD_SHAPE, "conv2d_wgrad: ld smaller than channel count");
    SAUNET_CHECK_ARG((d->q_scale == nullptr) == (d->q_shift == nullptr), SAUNET_ERR_BAD_SHAPE, "conv2d_wgrad: q_scale/q_shift mismatch");
    if (conv_wgrad_pw_eligible(d) && !SAUNET_ENV_FLAG("SAUNET_NO_WGRAD_PW")) return conv_wgrad_pw(d, (cudaStream_t)stream);
    if (conv_wgrad_skinny_eligible(d)) return conv_wgrad_skinny(d, (cudaStream_t)stream);
    if (conv_wgrad_halo_eligible(d) && !SAUNET_ENV_FLAG("SAUNET_NO_WGRAD_HALO")) return conv_wgrad_halo(d, (cudaStream_t)stream);
    if (conv_wgrad_tc_eligible(d)) return conv_wgrad_tc(d, (cudaStream_t)stream);
    return conv_wgrad_simt(d, (cudaStream_t)stream);
}
