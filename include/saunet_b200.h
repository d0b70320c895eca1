/* saunet_b200.h -- C ABI of libsaunet_b200.so (sm_100a).
 *
 * The reference (sunjesse/shape-attentive-unet) is pure Python/PyTorch and has
 * NO FFI/plugin layer for this path: its hot path calls
 * stock torch ops.  Each entry point below therefore names the reference
 * call site(s) whose ATen op it replaces (paths relative to /root/reference).
 * A reference maintainer binds these with ctypes.
 *
 * Conventions
 *   - every function returns int: 0 = ok, negative = SAUNET_ERR_*;
 *     saunet_last_error() gives the thread-local message; nothing throws/aborts.
 *   - all pointers are DEVICE pointers unless the name says host; the library
 *     never allocates, frees or retains caller memory; scratch is caller-owned.
 *   - activations are fp32 NHWC ("channels_last"): element (b,y,x,c) lives at
 *     ptr[((b*H + y)*W + x)*ld + c]; `ld` >= C lets a call address a channel
 *     slice of a wider (concatenated) buffer without a copy.
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it.
 *   - no CPU fallback exists anywhere in this library.
 */
#ifndef SAUNET_B200_H
#define SAUNET_B200_H
#ifdef __cplusplus
extern "C" {
#endif

#define SAUNET_OK 0
#define SAUNET_ERR_BAD_SHAPE (-1)
#define SAUNET_ERR_BAD_DTYPE (-2)
#define SAUNET_ERR_BAD_ALIGN (-3)
#define SAUNET_ERR_WORKSPACE (-4)
#define SAUNET_ERR_CUDA (-5)

#define SAUNET_ACT_NONE 0
#define SAUNET_ACT_RELU 1
#define SAUNET_ACT_SIGMOID 2

int saunet_version(void);
const char* saunet_last_error(void);
/* number of kernels this library has launched in the calling process (bench.py's gpu_launches) */
long long saunet_launch_count(void);
/* kernel family launched by this thread's most recent call (e.g. "conv_wgrad_halo_kernel"): bench.py's roofline attribution */
const char* saunet_last_kernel(void);

/* ---- implicit-GEMM convolution --------------------------------------------------------------
 * One descriptor drives forward, data-gradient and (with saunet_conv2d_wgrad) weight-gradient of
 * every conv-like op on the path:
 *   nn.Conv2d 1x1/3x3/7x7  (torchvision densenet.py:31-133,170; models/models.py:118-123,280-301,324;
 *                           models/resnet.py:24-27; models/attention_blocks.py:31-36,149-150,215-218)
 *   nn.ConvTranspose2d k4 s2 p1 (models/attention_blocks.py:179-183; models/models.py:211) as 4 phases
 *   F.conv2d in GatedSpatialConv2d (models/GSConv.py:56-57)
 * GEMM view: rows = output-grid pixels (b,i,j), i<Hg, j<Wg; K = KH*KW*Cin ordered (ky,kx,c); N = Cout.
 *   input pixel for tap (ky,kx) : (i*sy + ky + offy, j*sx + kx + offx), zero outside the image
 *   output pixel                : (i*osy + oy0,     j*osx + ox0)
 * Optional fused prologue on the gathered input (the pre-activation BatchNorm+ReLU of DenseNet,
 * densenet.py:47-50): a = relu?(x*in_scale[c] + in_shift[c]), applied BEFORE zero padding.
 * Optional fused epilogue: + bias[n]; per-output-channel sum / sum-of-squares accumulated (double
 * atomics) into stat_sum/stat_sumsq for the BatchNorm that follows; * (row_scale[p] + row_scale_add)
 * (the GSConv gate, GSConv.py:55); activation; accumulate into y.
 */
typedef struct saunet_conv_desc {
    const float* x; int x_ld; int B, Hin, Win, Cin;
    const float* w;              /* packed [KH*KW*Cin][Cout] (see saunet_pack_weights) */
    int Cout, KH, KW;
    int Hg, Wg, sy, sx, offy, offx;
    float* y; int y_ld; int Hout, Wout, osy, osx, oy0, ox0;
    const float* in_scale; const float* in_shift; int in_relu;
    const float* bias;
    const float* row_scale; float row_scale_add;
    int act;                     /* SAUNET_ACT_* */
    int accumulate;              /* y += result instead of y = result */
    double* stat_sum; double* stat_sumsq;
    /* tensor-core path (conv_tc.cu): weights tiled by saunet_pack_weights_tc for N tile tc_bn; tc_passes 3 = 3xTF32
     * (fp32-class accuracy), 1 = single-pass TF32, 16 = bf16 operands (kind::f16, fp32 accumulate; activations stay
     * fp32 in memory and are rounded to bf16 on their way into shared memory).  NULL w_tc (or an ineligible geometry: Cin % 4 != 0, unaligned
     * x) selects the exact-fp32 FFMA kernel, which reads `w`. */
    const float* w_tc; int tc_bn; int tc_passes;
    /* 1: w_tc is the chunk-major PADDED image of saunet_pack_weights_tc_cm (k-blocks = (32-channel chunk, tap), last
     * chunk zero-padded): what the TMA-fed 3x3 kernel (conv_halo_tma.cu) reads when Cin % 32 != 0 */
    int tc_cm;
    /* fused BatchNorm(+ReLU)-backward epilogue (1x1 data gradients, conv_pw_t.cu): the GEMM result is
     * da = d loss / d act(bn(x)) for the tensor epi_x (same pixels, channel n of epi_x <-> output channel n):
     *   g = da * [epi_relu ? epi_scale*x + epi_shift > 0 : 1];  y (+)= epi_scale * g;
     *   stat_sum[n] += sum g;  stat_sumsq[n] += sum g * (x - epi_mean[n])
     * i.e. the data-dependent term of the BatchNorm gradient; the mean terms follow from the sums
     * (saunet_bn_fused_finish, saunet_bn_fixup).  NULL epi_x = plain epilogue. */
    const float* epi_x; int epi_x_ld; const float* epi_scale; const float* epi_shift; const float* epi_mean; int epi_relu;
} saunet_conv_desc;

int saunet_conv2d_fwd(const saunet_conv_desc* d, void* stream);

/* weight gradient: dw[(ky,kx,cb)][ca] += sum_pixels P[pix][ca] * Q[gather(pix,ky,kx)][cb]
 * (regular conv: P = dY, Q = x; conv-transpose: P = x, Q = dY gathered with stride 2).
 * Q may carry the same prologue as the forward (recomputes relu(bn(x)) instead of storing it).
 * dw must be zeroed by the caller (accumulated with fp32 atomics over pixel splits). */
typedef struct saunet_wgrad_desc {
    const float* p; int p_ld; int Ca;
    const float* q; int q_ld; int Cb; int B, Hq, Wq;
    int KH, KW, Hg, Wg, sy, sx, offy, offx;
    const float* q_scale; const float* q_shift; int q_relu;
    float* dw;                   /* packed [KH*KW*Cb][Ca] */
    int precision;               /* 0: exact fp32 FFMA;  1: tcgen05 3xTF32 when the geometry allows;  2: tcgen05 single-pass TF32 */
} saunet_wgrad_desc;
int saunet_conv2d_wgrad(const saunet_wgrad_desc* d, void* stream);

/* weight (un)packing between PyTorch parameter layout w[A][Bc][KH][KW] and the packed GEMM layouts.
 *  mode 0: packed[(t,b)][a]       = w[a][b][t]                      (conv fwd, convT dgrad)
 *  mode 1: packed[(t,a)][b]       = w[a][b][KH*KW-1-t]              (stride-1 conv dgrad)
 *  mode 2: packed[ph][(ty,tx,a)][b] = w[a][b][3-pa-2ty][3-pb-2tx], ph = pa*2+pb  (convT 4x4 s2 p1 fwd phases)
 * unpack: w_grad[a][b][t] (+)= packed[(t,b)][a]  (mode 0 only). */
int saunet_pack_weights(const float* w, float* packed, int A, int Bc, int KH, int KW, int mode, void* stream);
/* tensor-core weight tiling: kn = packed [K][N] (any mode above, one phase for mode 2) ->
 * out[n_tile][k_block(32)][hi,lo][BN][32] in the UMMA K-major SWIZZLE_128B shared-memory image, tf32-split
 * (passes 3; passes 1: hi image only; passes 16: one bf16 image of 64-byte rows, SWIZZLE_64B, half the floats);
 * K = taps*Cin, where Cin is the channel count of the tensor the conv GATHERS (so dgrad passes Cout). */
int saunet_tc_tile_n(int Cout);
long long saunet_tc_packed_floats(int K, int N, int BN, int passes);
int saunet_tc_chunk_major(int taps, int Cin);   /* 1: K-blocks ordered (32-channel chunk, tap) for L1 reuse across taps */
int saunet_pack_weights_tc(const float* kn, int taps, int Cin, int N, int BN, int passes, float* out, void* stream);
long long saunet_tc_packed_floats_cm(int taps, int Cin, int N, int BN, int passes);
int saunet_pack_weights_tc_cm(const float* kn, int taps, int Cin, int N, int BN, int passes, float* out, void* stream);
int saunet_unpack_wgrad(const float* packed, float* wgrad, int A, int Bc, int KH, int KW, int accumulate, void* stream);
/* the same for every conv weight of a model in one launch (flat gradient arena, saunet_b200.parallel.GradArena):
 * grad_base[grad_off + ((a*Bc + b)*T + t)] += packed_base[packed_off + ((t*Bc + b)*A + a)], then that packed
 * element is reset to 0 (the images stay all-zero between backward passes); table lives in device
 * memory, `first` = exclusive prefix sum of A*Bc*T over the entries, total = their sum. */
typedef struct saunet_unpack_entry {
    long long first, packed_off, grad_off;
    int A, Bc, T, pad_;
} saunet_unpack_entry;
int saunet_unpack_wgrad_multi(const saunet_unpack_entry* table, int n, long long total, float* packed_base,
                              float* grad_base, void* stream);   /* (clears every packed element it consumes) */

/* ---- batch norm (nn.BatchNorm2d / SynchronizedBatchNorm2d outside DataParallel == F.batch_norm;
 *      lib/nn/modules/batchnorm.py:58-61; Appendix A of SURVEY.md) ---------------------------------- */
/* per-channel sum and sum of squares over npix pixels, accumulated into double[C] each (caller zeroes;
 * sum and sumsq must belong to one allocation with sumsq >= sum + C) */
int saunet_channel_stats(const float* x, int ld, int C, long long npix, double* sum, double* sumsq, void* stream);
/* dst[i] += (float)src[i]  -- folds an fp64 reduction (e.g. a conv bias gradient = column sums) into an fp32 grad */
int saunet_add_d2f(const double* src, float* dst, int n, void* stream);
/* state = float[4][C]: scale, shift, mean, invstd.  training: batch stats from sum/sumsq/count and
 * running-stat update (momentum; unbiased var);  eval: from running stats. */
int saunet_bn_finalize(const double* sum, const double* sumsq, double count, const float* gamma, const float* beta,
                       float* running_mean, float* running_var, float momentum, float eps, int training, int C,
                       float* state, void* stream);
/* y = act(x*scale[c] + shift[c] (+ residual)) */
int saunet_affine_act(const float* x, int x_ld, const float* scale, const float* shift, const float* residual,
                      int r_ld, float* y, int y_ld, int C, long long npix, int act, void* stream);
/* g = dy * act'(.)  where the ReLU mask comes from `out` (post-activation tensor) if given, else from
 * recomputing x*scale+shift > 0;  red[0][c] += sum g, red[1][c] += sum g*xhat  (double[2][C], caller zeroes) */
int saunet_bn_bwd_reduce(const float* dy, int dy_ld, const float* x, int x_ld, const float* out, int out_ld,
                         const float* state, int C, long long npix, int act, double* red, void* stream);
/* dx (+)= gamma*invstd*(g - sum_g/n - xhat*sum_gx/n)  [training]  or  g*scale [eval];
 * dres (+)= g if dres != NULL;  dgamma[c] += sum_gx, dbeta[c] += sum_g (fp32 parameter grads) */
int saunet_bn_bwd_apply(const float* dy, int dy_ld, const float* x, int x_ld, const float* out, int out_ld,
                        const float* state, const float* gamma, const double* red, int C, long long npix, int act,
                        int training, float* dx, int dx_ld, int dx_acc, float* dres, int dres_ld, int dres_acc,
                        float* dgamma, float* dbeta, void* stream);

/* Fused BatchNorm backward, second half (first half = the epi_x epilogue of saunet_conv2d_fwd): from
 * sums[0][c] = sum g, sums[1][c] = sum g*(x-mean) of ONE BatchNorm application over channels [0,C):
 *   dbeta[c] += sum g;  dgamma[c] += invstd*sum g(x-mean);
 *   training: ab[0][c] += scale*(sum g)/n - B*mean,  ab[1][c] += B,  B = scale*invstd^2*sum g(x-mean)/n
 * so that the complete input gradient is  dx = (sum over applications of scale*g)  -  (ab[0][c] + ab[1][c]*x).
 * DenseNet: a feature channel feeds the norm1 of EVERY later layer (densenet.py:47-50); their ab terms add up and ONE
 * fix-up pass per 32-channel slice replaces a reduce + an apply pass over all input channels per layer.
 * state = [scale, shift, mean, invstd][C] (saunet_bn_finalize); ab = double[2][ab_ld]. */
int saunet_bn_fused_finish(const double* sums, const float* state, double count, int training, int C, float* dgamma,
                           float* dbeta, double* ab, int ab_ld, void* stream);
/* g[p][c] -= ab[0][c] + ab[1][c] * x[p][c]   for c in [0,C)  (ab already offset to the first channel) */
int saunet_bn_fixup(float* g, int g_ld, const float* x, int x_ld, const double* ab, int ab_ld, int C, long long npix,
                    void* stream);

/* ---- pointwise / resampling ------------------------------------------------------------------- */
/* F.interpolate(mode='bilinear', align_corners=True) models/models.py:337-389 */
int saunet_bilinear_fwd(const float* x, int x_ld, int B, int Hin, int Win, int C, float* y, int y_ld, int Hout,
                        int Wout, void* stream);
int saunet_bilinear_bwd(const float* dy, int dy_ld, int B, int Hin, int Win, int C, float* dx, int dx_ld, int Hout,
                        int Wout, int accumulate, void* stream);
/* nn.AvgPool2d(2,2) densenet.py:133 ; nn.MaxPool2d(2,2) models/models.py:269,376 */
int saunet_avgpool2_fwd(const float* x, int x_ld, int B, int Hin, int Win, int C, float* y, int y_ld, void* stream);
int saunet_avgpool2_bwd(const float* dy, int dy_ld, int B, int Hin, int Win, int C, float* dx, int dx_ld, int accumulate, void* stream);
int saunet_maxpool2_fwd(const float* x, int x_ld, int B, int Hin, int Win, int C, float* y, int y_ld, void* stream);
int saunet_maxpool2_bwd(const float* dy, int dy_ld, const float* x, int x_ld, int B, int Hin, int Win, int C, float* dx,
                        int dx_ld, int accumulate, void* stream);
/* nn.AdaptiveAvgPool2d(1) models/attention_blocks.py:32,51 : y[b][c] = mean over HW */
int saunet_gap_fwd(const float* x, int x_ld, int B, long long HW, int C, float* y, void* stream);
int saunet_gap_bwd(const float* dy, int B, long long HW, int C, float* dx, int dx_ld, int accumulate, void* stream);
/* dz = dy * act'(y) from the activation OUTPUT y */
int saunet_act_bwd(const float* dy, int dy_ld, const float* y, int y_ld, float* dz, int dz_ld, int C, long long npix, int act, void* stream);
/* DualAttBlock combine (attention_blocks.py:57,237): out = fused * (1 + S[p]) * cvec[b][c] */
int saunet_dualatt_combine_fwd(const float* fused, int f_ld, const float* S, const float* cvec, int B, long long HW,
                               int C, float* out, int o_ld, void* stream);
/* dfused (+)= dout*(1+S)*c ; dS[p] (=) sum_c dout*fused*c ; dcvec[b][c] += sum_p dout*fused*(1+S) (caller zeroes dcvec) */
int saunet_dualatt_combine_bwd(const float* dout, int do_ld, const float* fused, int f_ld, const float* S,
                               const float* cvec, int B, long long HW, int C, float* dfused, int df_ld, int df_acc,
                               float* dS, float* dcvec, void* stream);
/* GSConv gate backward (GSConv.py:55-57 with out = (alpha+1)*conv(x)):
 * du = dout*(alpha+1) ; dalpha[p] (+)= sum_c dout*out/(alpha+1) */
int saunet_rowscale_bwd(const float* dout, int do_ld, const float* out, int o_ld, const float* alpha, int C,
                        long long npix, float* du, int du_ld, float* dalpha, int da_acc, void* stream);
/* strided channel-slice copy / accumulate: dst[p][0:C] (+)= src[p][0:C] */
int saunet_copy_slice(const float* src, int s_ld, float* dst, int d_ld, int C, long long npix, int accumulate, void* stream);
/* layout conversion at the module boundary (NCHW <-> NHWC) */
int saunet_nchw_to_nhwc(const float* src, float* dst, int dst_ld, int B, int C, long long HW, void* stream);
int saunet_nhwc_to_nchw(const float* src, int src_ld, float* dst, int B, int C, long long HW, void* stream);

/* ---- loss: loss.py:51-88 (dice_loss), :149-159 (DualLoss.forward) --------------------------------
 * logits NHWC [npix][C] (C <= 8), edge prob [npix], seg target int64 [npix], edge target float [npix].
 * acc = double[2 + 2*C + 1]: [0]=sum w*nll, [1]=sum w, [2..2+C)=I_c, [2+C..2+2C)=Card_c, [2+2C]=sum bce (caller zeroes)
 * parts: bit0 = dice, bit1 = weighted CE, bit2 = edge BCE -- the terms summed into loss[0] and differentiated
 * (DualLoss = 7; loss.dice_loss alone = 1).
 * finalize writes loss[0]=total of the selected parts, [1]=dice, [2]=ce, [3]=bce (float).
 * Labels outside [0,C) are skipped by every sum (never used as an index) and counted in counts[2+3*7].
 * counts (optional, int[2 + 3*7 + 1], caller zeroes): the training-branch metrics of SegmentationModule
 * (models/models.py:51-74,92) taken from the same pass: pred = argmax(round(softmax(logits))); [0] = |label>=1 & pred==label|,
 * [1] = |label>=1|, class i>=1: [2+3(i-1)] = |label==i & pred==i|, [+1] = |label==i|, [+2] = |pred==i|; finalize then also
 * writes loss[4] = pixel accuracy and loss[4+i] = Jaccard of class i (loss must then hold 4 + C floats). */
int saunet_dual_loss_fwd(const float* logits, int l_ld, const float* edge, const long long* seg_t, const float* edge_t,
                         long long npix, int C, const float* class_w, int parts, double* acc, float* loss, int* counts,
                         void* stream);
int saunet_dual_loss_bwd(const float* logits, int l_ld, const float* edge, const long long* seg_t, const float* edge_t,
                         long long npix, int C, const float* class_w, const double* acc, const float* dloss,
                         float* dlogits, int dl_ld, float* dedge, int parts, void* stream);

/* ---- optimizer: radam.py:15-78 (RAdam), torch.optim.SGD / Adam as built by train.py:188-207 over the two groups of
 *      train.py:166-185 -- one fused pass over the flat parameter arena instead of a Python loop over ~700 tensors.
 * w, g, m, v: flat fp32 arrays of n elements (n % 4 == 0; every tensor starts at a multiple of 4);
 * seg_table: device array of nseg {long long begin; float lr; float weight_decay} sorted by begin (one per tensor);
 * step_counter: device int, number of steps taken so far (read, then incremented on the stream): bias corrections and
 * the RAdam rectification term are computed from it in the kernel, so the call is CUDA-graph capturable.
 * kind 0 = SGD(momentum, nesterov=False), 1 = Adam, 2 = RAdam (weight decay applied as radam.py:67-68). */
int saunet_optimizer_step(int kind, float* w, const float* g, float* m, float* v, long long n, const void* seg_table,
                          int nseg, int* step_counter, float beta1, float beta2, float eps, float momentum, void* stream);

/* ---- volume inference (test_and_pack.py:98-137) and loader-side data prep (data/ac17_dataloader.py:231-258) ----
 * argmax over the C logits of every pixel (first maximum on ties, as torch.max) -> uint8 label map */
int saunet_argmax_u8(const float* logits, int ld, int C, long long npix, unsigned char* out, void* stream);
/* edge ground truth of a label map (int64 [B][H][W], classes 1..num_classes): out[p] = 1.0 iff a pixel within Euclidean
 * distance `radius` (2 in the reference) carries a different label, pixels just outside the image counting as 0 --
 * identical to the reference's two-distance-transforms-per-class construction (mask_to_edges) */
int saunet_edge_gt(const long long* seg, int B, int H, int W, int radius, int num_classes, float* out, void* stream);

/* ---- Canny fusion: models/models.py:358-364 (np.mean(axis=1).astype(uint8) + cv2.Canny(im,10,100)) ----
 * x is the fp32 image, NCHW [B][C][H][W]; out is float [B][H][W] in {0,255}.
 * workspace: saunet_canny_workspace_bytes(B,H,W) bytes. */
long long saunet_canny_workspace_bytes(int B, int H, int W);
int saunet_canny_fwd(const float* x_nchw, int B, int C, int H, int W, int low, int high, float* out, void* workspace,
                     long long workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif
