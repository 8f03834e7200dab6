/*
 * mvdseg.h -- C ABI of libmvdseg.so: the hand-written sm_100a kernels behind the nnU-Net v2 3d_fullres
 * training step of JaronTu/Multimodal_MVD_Seg (PlainConvUNet fwd/bwd, deep-supervision Dice+CE, mutual-distillation
 * KL, soft-skeleton clDice, clip + SGD-nesterov).
 *
 * The reference has no FFI of its own: it reaches the GPU through torch.nn modules / ATen.  Each
 * entry point below therefore names the reference call (file:line under nnUNet/nnunetv2/) whose device work it
 * replaces.  INTEGRATION.md shows the ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *   - plain C: raw device pointers, ints, floats; no torch / C++ types.  The caller owns every buffer (outputs and
 *     workspaces included); the library allocates nothing persistent.
 *   - activations are bf16 "NDHWC": logical [B][D][H][W][C], channel contiguous, voxel pitch `ld` elements
 *     (ld >= C; ld > C addresses a channel slice of a wider buffer, which is how the decoder's torch.cat,
 *     training/my_network/UNetDecoder.py:107, is made copy-free).  Voxels are dense: sample pitch = D*H*W*ld.
 *   - every call is asynchronous on `stream`, performs no host synchronisation and no allocation, and is CUDA-graph
 *     capturable.  Device is taken from the current context of the calling thread (one process per GPU).
 *   - return 0 on success, a negative mvd_status otherwise; text via mvd_last_error() (thread-local).
 */
#ifndef MVDSEG_H_
#define MVDSEG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* mvd_stream_t; /* cudaStream_t */

enum mvd_status {
  MVD_OK = 0,
  MVD_ERR_INVALID = -1,     /* bad argument / unsupported shape */
  MVD_ERR_CUDA = -2,        /* CUDA runtime / driver error      */
  MVD_ERR_UNSUPPORTED = -3  /* no kernel for this case          */
};

/* ---- library ---------------------------------------------------------------------------------------------- */
int mvd_version(void);
const char* mvd_last_error(void);
/* number of kernels this library has launched since load / last reset (bench.py's gpu_launches) */
unsigned long long mvd_launch_count(void);
void mvd_reset_launch_count(void);
/* number of convolution calls with algo == 0 (auto) that were NOT covered by a tcgen05 kernel and ran on the CUDA-core
 * tiles instead (csrc/conv_generic.cu) since load / last reset: bench.py prints it, the benchmark configurations need 0 */
unsigned long long mvd_fallback_count(void);
void mvd_reset_fallback_count(void);
int mvd_shutdown(void);
/* measurement aid: one thread busy-waits `cycles` SM clocks on `stream` (not counted in mvd_launch_count) */
int mvd_spin(long long cycles, mvd_stream_t stream);

/* ---- layout at the module edge ---------------------------------------------------------------------------- */
/* data.to(device) then network(data): fp32 NCDHW batch -> bf16 NDHWC (nnUNetTrainer.py:895,907).  src_batch_stride
 * (elements; 0 = dense C*V) lets a channel slice of a wider batch be read in place: the per-modality split data[:, 0:1] /
 * data[:, 1:2] of the mutual-distillation step (selfattnNet.py:588-589) */
int mvd_ncdhw_f32_to_ndhwc_bf16(const float* src, long long src_batch_stride, void* dst, int B, int C, long long V,
                                int ld_dst, mvd_stream_t stream);
/* logits back to fp32 NCDHW for callers that want the reference's memory format */
int mvd_ndhwc_bf16_to_ncdhw_f32(const void* src, int ld_src, float* dst, int B, int C, long long V,
                                mvd_stream_t stream);

/* ---- convolution ------------------------------------------------------------------------------------------ */
/* Replaces torch.nn.Conv3d / ConvTranspose3d fprop, dgrad, wgrad as instantiated at
 * utilities/get_network_from_plans.py:70-83 and training/my_network/UNetDecoder.py:55-65.
 * One geometry struct serves all six entry points; (Di,Hi,Wi,Cin) is always the conv INPUT side
 * (for a transposed conv the "conv" is its adjoint, i.e. Cin = the ConvTranspose3d's out_channels). */
/* InstanceNorm + LeakyReLU that sits in FRONT of the tensor a data gradient is produced for (the block whose output the
 * conv consumed): with it mvd_conv3d_dgrad also returns that block's backward statistics, so the separate pass
 * mvd_inorm_lrelu_bwd_stats over the gradient it just wrote is not needed (get_network_from_plans.py:41-44: conv -> norm ->
 * nonlin blocks back to back). */
typedef struct {
  const void* y; int ldy;     /* bf16 raw output of the conv in front of that norm, dense voxel order, pitch ldy      */
  const double* stats;        /* its forward statistics [B][Cin][2] (sum, sum of squares)                             */
  const float* gamma;         /* affine weight / bias, fp32 [Cin]; may be NULL (1 / 0)                                */
  const float* beta;
  float eps, slope;
  double* bstats;             /* out, accumulated (caller zeroes): [B][Cin][2] exactly as mvd_inorm_lrelu_bwd_stats   */
} mvd_norm_bwd_stats_args;

typedef struct {
  int B;
  int Di, Hi, Wi, Cin;   /* conv input  [B,Di,Hi,Wi,Cin]  */
  int Do, Ho, Wo, Cout;  /* conv output [B,Do,Ho,Wo,Cout] */
  int kd, kh, kw;        /* kernel                                   */
  int sd, sh, sw;        /* stride                                   */
  int pd, ph, pw;        /* zero padding                             */
  const void* x;   int ldx;   /* bf16 conv-input-side activation (fprop in, dgrad out, wgrad in)            */
  const void* y;   int ldy;   /* bf16 conv-output-side activation (fprop out, dgrad in, wgrad in)           */
  const void* w;              /* bf16 packed weights, see mvd_pack_conv_weights                             */
  const float* bias;          /* fp32 [Cout] (fprop) / [Cin] (transposed fprop); may be NULL                */
  double* stats;              /* optional [B][C][2] running (sum, sum of squares) of the bf16-rounded output,
                                 accumulated (caller zeroes): InstanceNorm statistics (fprop, C = Cout); for
                                 dgrad the sums of the produced gradient (C = Cin, ignored with accumulate)  */
  float* dw;                  /* wgrad out, fp32 in torch layout [Cout][Cin][kd][kh][kw]                    */
  float* dbias;               /* wgrad out, fp32 [Cout]; may be NULL                                        */
  void* workspace; size_t workspace_bytes; /* scratch (wgrad partials; fprop / dgrad split-K partials of small layers:
                                              optional there, without it the unsplit kernel runs); see
                                              mvd_conv3d_workspace_bytes */
  int algo;                   /* 0 auto, 1 CUDA-core tiles, 2 tcgen05 implicit GEMM                          */
  int accumulate;             /* dgrad: add into the output instead of overwriting it                       */
  const mvd_norm_bwd_stats_args* norm_bwd; /* dgrad only, optional: also produce the backward statistics of the
                                 InstanceNorm + LeakyReLU in front of x from the FINAL values written to x (after
                                 accumulate).  Computed in the conv epilogue where mvd_conv3d_dgrad_fuses_norm_bwd says so,
                                 by a pass over x otherwise. */
} mvd_conv3d_args;

/* Sliding-window inference accumulators (inference/predict_from_raw_data.py:703-712):
 *   acc[k][z0+z][y0+y][x0+x] += pred[fz][fy][fx][k] * scale * g[z][y][x];  npred[...] += g  (g = 1 when gaussian ==
 *   NULL; npred may be NULL).  pred: bf16 NDHWC tile [d][h][w][K] with voxel pitch ldp; acc fp32 [K][D][H][W]; npred
 *   fp32.  flip_mask (bit 0: z, bit 1: y, bit 2: x) reads the tile mirrored (fz = d-1-z ...): the flip-back of a
 *   mirrored test-time-augmentation pass (predict_from_raw_data.py:562-589), so that every pass is accumulated in fp32
 *   with scale = 1/2^n_axes and the averaged prediction is never rounded to 16 bits. */
int mvd_sw_accumulate(const void* pred, int ldp, const float* gaussian, float scale, float* acc, float* npred, int K,
                      int d, int h, int w, int D, int H, int W, int z0, int y0, int x0, int flip_mask,
                      mvd_stream_t stream);
int mvd_sw_finalize(float* acc, const float* npred, int K, long long vol, mvd_stream_t stream);   /* acc /= npred */
/* ---- training-batch augmentation on the GPU ------------------------------------------------------
 * Replaces the batchgenerators transform chain of nnUNetTrainer.get_training_transforms (MVDTrainer.py:700-765; the
 * transforms themselves live in the third-party package batchgenerators, absent from the reference tree).  Volumes are
 * fp32 planes [N][D][H][W], N = B * C.  The random draws are explicit per-plane / per-sample parameter arrays in DEVICE
 * memory; multimodal_mvd_seg_b200/augment.py samples them with the reference's distributions. */
/* cubic B-spline coefficients in place (scipy.ndimage.spline_filter, order 3, mode 'mirror'); apply: [N] flags or NULL */
int mvd_aug_spline_prefilter(float* vol, int N, int D, int H, int W, const unsigned char* apply, mvd_stream_t stream);
/* SpatialTransform (MVDTrainer.py:700-711): dst[b][c] = src[b][c] sampled at  M_b * (idx - (out - 1) / 2) + in / 2 - 0.5
 * (augment_spatial's zero-centred mesh, random_crop = False).  mat: [B][9] row-major; mode: [B], 0 = centre crop without
 * interpolation (no rotation / scaling drawn), 1 = interpolate.  seg_labels == 0: images, order 3 (src holds the spline
 * coefficients of mvd_aug_spline_prefilter for the samples with mode 1) or 1, constant border cval.  seg_labels > 0: the
 * segmentation rule of interpolate_img(is_seg = True, order = 1): label = the largest c in [1, seg_labels) whose linearly
 * interpolated mask is >= 0.5, else 0 (RemoveLabelTransform(-1, 0), MVDTrainer.py:738, folded in). */
int mvd_aug_spatial(const float* src, int B, int C, int Di, int Hi, int Wi, float* dst, int D, int H, int W,
                    const float* mat, const int* mode, int order, float cval, int seg_labels, mvd_stream_t stream);
/* GaussianNoiseTransform (:716): x += N(0, sigma[plane]); sigma 0 skips the plane; counter-based generator */
int mvd_aug_gaussian_noise(float* x, long long V, int N, const float* sigma, unsigned long long seed, mvd_stream_t stream);
/* GaussianBlurTransform (:717-718): scipy.ndimage.gaussian_filter (truncate 4, 'reflect') with sigma[plane]; 0 skips */
int mvd_aug_gaussian_blur(float* x, float* tmp, int N, int D, int H, int W, const float* sigma, mvd_stream_t stream);
/* out[N][4] doubles (sum, sum of squares, min, max), accumulated: the caller initialises (0, 0, +inf, -inf) */
int mvd_aug_plane_stats(const float* x, long long V, int N, double* out, mvd_stream_t stream);
/* op 0: x *= a[p] (BrightnessMultiplicativeTransform, :719).  op 1: ContrastAugmentationTransform (:720, preserve_range):
 * x = clip((x - mean) a[p] + mean, min, max) with stats0 of x; a = 0 skips.  op 2 / 3: GammaTransform (:726-727,
 * retain_stats): op 2 maps s x (s = -1 when invert) through ((. - min) / (range + 1e-7)) ^ a[p] * range + min, op 3 restores
 * mean / std (stats0 = before op 2, stats1 = after op 2) and undoes the inversion; a = 0 skips. */
int mvd_aug_intensity(float* x, long long V, int N, int op, const float* a, const double* stats0, const double* stats1,
                      int invert, mvd_stream_t stream);
/* SimulateLowResolutionTransform (:721-725): per plane with tshape[p] != 0, nearest-neighbour down-sampling to tshape[p]
 * (= round(shape * zoom)) and cubic up-sampling back (skimage resize(mode='edge', anti_aliasing=False) = scipy zoom(
 * grid_mode=True, mode='nearest'), clipped to the range of the low-resolution volume).  buf: N x buf_stride floats of
 * scratch, buf_stride >= (D+24)(H+24)(W+24); minmax: [N][2] doubles initialised (+inf, -inf). */
int mvd_aug_simulate_lowres(float* x, int N, int D, int H, int W, const int* tshape, float* buf, long long buf_stride,
                            double* minmax, mvd_stream_t stream);
/* MirrorTransform (:729-730): flips[B][3] (d, h, w) per sample, out of place */
int mvd_aug_mirror(const float* src, float* dst, int B, int C, int D, int H, int W, const unsigned char* flips,
                   mvd_stream_t stream);

/* Deep-supervision targets on the GPU.  Replaces DownsampleSegForDSTransform2.__call__
 * (training/data_augmentation/custom_transforms/deep_supervision_donwsampling.py:27-55; batchgenerators'
 * resize_segmentation with order 0): nearest-neighbour with pixel-centre alignment, src = floor((o + 0.5) * I / O) per
 * axis.  seg: fp32 [BC][Di][Hi][Wi] (BC = batch x seg channels); dst[s]: fp32 [BC][Do_s][Ho_s][Wo_s] with
 * out_dhw = {Do_0, Ho_0, Wo_0, Do_1, ...} (HOST arrays, n_scales <= 8); all scales in one launch. */
int mvd_downsample_seg_nearest(const float* seg, int BC, int Di, int Hi, int Wi, int n_scales, float* const* dst,
                               const int* out_dhw, mvd_stream_t stream);
/* The stem (first conv: Cin = 1 or 2 modalities -> 32 features, 3x3x3, stride 1, pad 1) as a tensor-core GEMM whose
 * im2col tile is built in shared memory (csrc/stem_tc.cu); replaces nn.Conv3d(Cin, 32, 3, padding=1) of
 * get_network_from_plans.py:75-77 for the first block.
 *   x     : DENSE bf16 NDHWC input [B][D][H][W][Cin]
 *   wcol  : bf16 [32][KPAD], KPAD = 32*Cin, column k = tap*Cin + ci (tap = (kd*3 + kh)*3 + kw), zero padded
 *           (mvd_pack_conv_weights_multi writes it: mvd_pack_desc.stem_kpad)
 *   fprop : y (pitch ldy) = conv + bias (bf16), optional InstanceNorm sums stats [B][32][2] accumulated
 *   wgrad : dw fp32 [32][Cin][27] (the torch layout of the weight) = sum_v dy[v][co] * x[v + tap - 1][ci] (overwritten) */
int mvd_stem_conv_fprop(const void* x, int B, int D, int H, int W, int Cin, const void* wcol, const float* bias,
                        void* y, int ldy, double* stats, mvd_stream_t stream);
int mvd_stem_conv_wgrad(const void* x, int B, int D, int H, int W, int Cin, const void* dy, int lddy, float* dw,
                        mvd_stream_t stream);
/* pack fp32 torch-layout weights [Cout][Cin][kd][kh][kw] into the two bf16 GEMM layouts:
 *   w_fprop [tap][Cout][Cin]  (B operand of fprop:  N = Cout rows, K = Cin contiguous)
 *   w_dgrad [tap][Cin][Cout]  (B operand of dgrad:  N = Cin rows,  K = Cout contiguous)
 * either output may be NULL. tap = (kd*KH + kh)*KW + kw. */
typedef struct mvd_pack_desc {
  const float* w;         /* fp32 [Cout][Cin][taps]                                                         */
  void* w_fprop;          /* bf16 [tap][Cout][Cin] or NULL                                                  */
  void* w_dgrad;          /* bf16 [tap][Cin][Cout] or NULL                                                  */
  int Cout, Cin, taps;    /* taps <= 27                                                                     */
  int block_begin;        /* first thread block of this layer: running sum of mvd_pack_blocks() over the table */
  int stem_kpad;          /* 0: the two layouts above.  > 0 (Cin <= 16): w_fprop receives the stem layout instead,
                             bf16 [Cout][stem_kpad] with column k = tap*Cin + ci, zero padded (mvd_stem_conv_fprop)   */
  int reserved;
} mvd_pack_desc;
/* packs a whole table of layers (device memory, n entries, ascending block_begin) in one launch of total_blocks blocks */
int mvd_pack_conv_weights_multi(const mvd_pack_desc* descs_device, int n, int total_blocks, mvd_stream_t stream);
int mvd_pack_blocks(int Cout, int Cin);   /* thread blocks one layer occupies in the table */
/* The optimiser update of the conv weights fused with the refresh of their packed layouts: for every table entry
 * (p = pack.w, grad, momentum: fp32, same shape)  g = grad*coef + wd*p; buf = mom*buf + g; p -= lr*(g + mom*buf)  with
 * the clip coefficient of mvd_sgd_nesterov_clip, then pack.w_fprop / pack.w_dgrad are rewritten from the new weights.
 * The next forward pass needs no mvd_pack_conv_weights_multi (MVDTrainer.py:978-979 + autocast's weight casts, :894). */
typedef struct mvd_sgd_pack_desc {
  mvd_pack_desc pack;
  const float* grad;
  float* momentum;
} mvd_sgd_pack_desc;
int mvd_sgd_pack_conv_weights(const mvd_sgd_pack_desc* descs_device, int n, int total_blocks, const double* sqnorm,
                               float gscale, float max_norm, float lr, float weight_decay, float momentum,
                               mvd_stream_t stream);
/* Weight-gradient reduction mode.  0 (default): the splits of the voxel range reduce into one scratch with
 * red.global.add.v4.f32 (order-dependent last bits).  1: deterministic two-stage reduction -- every split stores its
 * partial dw into its own workspace slice and the slices are added in a fixed order (bit-reproducible; +1.2 % step time
 * at cfg-2).  Process-wide; changes the answer of mvd_conv3d_workspace_bytes(pass 2).  Env MVD_DETERMINISTIC=1 selects 1
 * at load.  (The split-K partials of small fprop / dgrad layers are always reduced in a fixed order.) */
int mvd_set_deterministic(int on);
int mvd_get_deterministic(void);
size_t mvd_conv3d_workspace_bytes(const mvd_conv3d_args* a, int pass /*0 fprop,1 dgrad,2 wgrad*/);
int mvd_conv3d_fprop(const mvd_conv3d_args* a, mvd_stream_t stream); /* y = conv(x, w) + bias  (w = w_fprop) */
int mvd_conv3d_dgrad(const mvd_conv3d_args* a, mvd_stream_t stream); /* x = conv^T(y, w)       (w = w_dgrad) */
int mvd_conv3d_wgrad(const mvd_conv3d_args* a, mvd_stream_t stream); /* dw = x (*) y, dbias = sum y          */
/* 1 when mvd_conv3d_dgrad(a) folds a->norm_bwd into its epilogue (no extra pass over the produced gradient) */
int mvd_conv3d_dgrad_fuses_norm_bwd(const mvd_conv3d_args* a);

/* stem (Cin = 1 or 2 input modalities): explicit im2col, X_col[v][tap*Cin + ci] (stride 1), zero-padded to Kpad columns
 * (a multiple of 8); the stem's fprop / wgrad then run as single-tap tensor-core GEMMs over X_col. */
int mvd_im2col_small(const void* x, int ldx, int B, int D, int H, int W, int Cin, int kd, int kh, int kw, int pd,
                     int ph, int pw, void* out, int Kpad, mvd_stream_t stream);

/* ---- InstanceNorm3d(affine, eps) + LeakyReLU -------------------------------------------------------------- */
/* Replaces nn.InstanceNorm3d + nn.LeakyReLU(inplace) of every ConvDropoutNormReLU block
 * (get_network_from_plans.py:41-44).  stats = [B][C][2] doubles (sum, sumsq) over the V voxels of each (b,c). */
int mvd_inorm_stats(const void* y, int ldy, int B, long long V, int C, double* stats, mvd_stream_t stream);
int mvd_inorm_lrelu_fwd(const void* y, int ldy, void* z, int ldz, const double* stats, const float* gamma,
                        const float* beta, int B, long long V, int C, float eps, float slope, mvd_stream_t stream);
/* bstats = [B][C][2] doubles: sum g', sum g'*xhat with g' = dz * lrelu'(.) ; caller zeroes */
int mvd_inorm_lrelu_bwd_stats(const void* dz, int lddz, const void* y, int ldy, const double* stats,
                              const float* gamma, const float* beta, int B, long long V, int C, float eps,
                              float slope, double* bstats, mvd_stream_t stream);
/* dy = gamma*rstd*(g' - mean(g') - xhat*mean(g' xhat)); dgamma/dbeta (fp32 [C]) written when non-NULL;
 * dsum (fp32 [C], caller zeroes, may be NULL) += per-channel sum of dy = bias gradient of the conv in front of the norm */
int mvd_inorm_lrelu_bwd_apply(const void* dz, int lddz, const void* y, int ldy, void* dy, int lddy,
                              const double* stats, const double* bstats, const float* gamma, const float* beta,
                              int B, long long V, int C, float eps, float slope, float* dgamma, float* dbeta,
                              float* dsum, mvd_stream_t stream);

/* InstanceNorm + LeakyReLU of the last decoder block folded into its only consumer, the 1x1x1 head (csrc/norm_head.cu;
 * C = 32 features, K = 4 classes -- mvd_inorm_lrelu_head_supported): the normalised activation and the head's data
 * gradient never touch HBM.
 *   fwd       : logits [B][V][4] (dense bf16) = head(lrelu(IN(y)))            y: raw conv output, stats as above
 *   bwd_stats : bstats (caller zeroes) from dz = dlogits W recomputed per voxel; head dw [4][32] / dbias [4] accumulated
 *   bwd_apply : dy (pitch lddy), dgamma / dbeta / dsum as mvd_inorm_lrelu_bwd_apply */
int mvd_inorm_lrelu_head_supported(int C, int K);
int mvd_inorm_lrelu_head_fwd(const void* y, int ldy, const double* stats, const float* gamma, const float* beta,
                             const float* w, const float* bias, void* logits, int B, long long V, int C, int K,
                             float eps, float slope, mvd_stream_t stream);
int mvd_inorm_lrelu_head_bwd_stats(const void* dlogits, const void* y, int ldy, const double* stats, const float* gamma,
                                   const float* beta, const float* w, int B, long long V, int C, int K, float eps,
                                   float slope, double* bstats, float* dw, float* dbias, mvd_stream_t stream);
int mvd_inorm_lrelu_head_bwd_apply(const void* dlogits, const void* y, int ldy, void* dy, int lddy, const double* stats,
                                   const double* bstats, const float* gamma, const float* beta, const float* w, int B,
                                   long long V, int C, int K, float eps, float slope, float* dgamma, float* dbeta,
                                   float* dsum, mvd_stream_t stream);

/* ---- 1x1x1 segmentation heads (UNetDecoder.py:67-70) ------------------------------------------------------- */
int mvd_head_fwd(const void* z, int ldz, const float* w /*[K][C] fp32*/, const float* bias /*[K]*/, void* logits,
                 int ldl, long long NV /*B*V*/, int C, int K, mvd_stream_t stream);
/* dz (bf16) = dlogits * w (accumulate_dz: += , the gradient another consumer of z already left there);
 * dw [K][C], dbias [K] fp32 accumulated (caller zeroes) */
int mvd_head_bwd(const void* dlogits, int ldl, const void* z, int ldz, const float* w, void* dz, int lddz,
                 float* dw, float* dbias, long long NV, int C, int K, int accumulate_dz, mvd_stream_t stream);

/* ---- deep-supervision Dice + CE (nnUNetTrainer.py:359-374; robust_ce_loss.py:12-16) ----------------------- */
/* acc = [B][C][3] doubles (intersect, sum_pred, sum_gt) followed by 1 double (sum of -log p[target]); caller zeroes.
 * logits bf16 NDHWC [B][V][C] pitch ld; target fp32 [B][V] holding class ids (the reference's float targets,
 * MVDTrainer.py:765).  C <= 8. */
int mvd_dice_ce_fwd(const void* logits, int ld, const float* target, int B, long long V, int C, double* acc,
                    mvd_stream_t stream);
/* loss_out[0] += weight * (w_ce*CE + w_dice*Dice); coef = [B][C][2] floats (a, e) kept for the backward.
 * batch_dice: sums over b first (MemoryEfficientSoftDiceLoss batch_dice=True). do_bg as in the reference. */
int mvd_dice_ce_finalize(const double* acc, int B, long long V, int C, float smooth, int do_bg, int batch_dice,
                         float w_ce, float w_dice, float weight, float* coef, float* loss_out, mvd_stream_t stream);
/* dlogits (bf16, pitch ldd) = gout[0] * weight * d(w_ce*CE + w_dice*Dice)/dlogits */
int mvd_dice_ce_bwd(const void* logits, int ld, const float* target, int B, long long V, int C, const float* coef,
                    float w_ce, float weight, const float* gout, void* dlogits, int ldd, mvd_stream_t stream);
/* The deep-supervision loss of a whole step in three launches (csrc/losses_multi.cu).  One segment per (network,
 * active scale); production shape only: C = 4, dense logits [B][V][4] (pitch 4), fp32 targets [B][V], V % 4 == 0,
 * 16-byte aligned pointers (anything else: the per-scale entry points above).  The table is a HOST array.
 *   fwd      : acc [n_seg][B*4*3 + 2] doubles (caller zeroes) <- per-(b,c) sums, the CE sum and the number of VALID voxels
 *              of every segment; when `counter` (device unsigned, zero before the first call; the kernel leaves it zero) is
 *              given, the last block to finish also writes coef [n_seg][B*4*2 + 1] ((A, E) per (b, c), then 1 / valid) and
 *              loss_out[0] = sum_seg weight * (w_ce CE + w_dice Dice)
 *   ignore label: a voxel whose target is outside [0, 4) is ignored -- it enters none of the Dice sums (loss_mask of
 *              MemoryEfficientSoftDiceLoss), the cross entropy is the mean over the valid voxels (ignore_index; 0 when there
 *              are none) and its dlogits are 0.  nnU-Net's ignore label is the id behind the last class, i.e. out of range.
 *   finalize : the same scalar algebra as its own launch (data-parallel batch_dice: the sums are all-reduced first)
 *   bwd      : dlogits of every segment = gout[0] * weight * d(w_ce CE + w_dice Dice)/dlogits; coef_scale multiplies
 *              the Dice coefficients (world size under AllGatherGrad, ddp_allgather.py:35-48; else 1) */
#define MVD_DICE_CE_MAX_SEGMENTS 16
typedef struct {
  const void* logits;    /* bf16 [B][V][4]                    */
  const float* target;   /* fp32 [B][V] class ids             */
  void* dlogits;         /* bf16 [B][V][4] (bwd only)         */
  long long V;           /* voxels per sample                 */
  float weight;          /* deep-supervision weight of the scale */
} mvd_dice_ce_segment;
int mvd_dice_ce_multi_fwd(const mvd_dice_ce_segment* segs, int n_seg, int B, int C, float smooth, int do_bg,
                          int batch_dice, float w_ce, float w_dice, double* acc, float* coef, float* loss_out,
                          unsigned* counter, mvd_stream_t stream);
int mvd_dice_ce_multi_finalize(const mvd_dice_ce_segment* segs, int n_seg, int B, int C, float smooth, int do_bg,
                               int batch_dice, float w_ce, float w_dice, const double* acc, float* coef,
                               float* loss_out, mvd_stream_t stream);
int mvd_dice_ce_multi_bwd(const mvd_dice_ce_segment* segs, int n_seg, int B, int C, const float* coef, float w_ce,
                          float coef_scale, const float* gout, mvd_stream_t stream);
/* validation_step's online tp/fp/fn of the argmax segmentation (nnUNetTrainer.py:973-1004): out = [C][3] doubles */
int mvd_argmax_tp_fp_fn(const void* logits, int ld, const float* target, int B, long long V, int C, double* out,
                        mvd_stream_t stream);

/* ---- mutual-distillation KL (training/loss/other_loss.py:51-64; call site MVDTrainer.py:897-899) ---------- */
/* C == 1 selects the reference's shape[1]==1 branch (the single logit against a constant zero logit, 2 classes).
 * loss_sum[0] (double, caller zeroes) accumulates sum p_t*(log p_t - log p_s); the host scales by T^2/numel. */
int mvd_kl_fwd(const void* ys, int lds, const void* yt, int ldt, long long NV, int C, float T, double* loss_sum,
               mvd_stream_t stream);
/* dys/dyt (bf16) = gout[0]*scale * dKL/dy ; scale = T^2/numel supplied by the caller; either may be NULL */
int mvd_kl_bwd(const void* ys, int lds, const void* yt, int ldt, long long NV, int C, float T, float scale,
               const float* gout, void* dys, int ldds, void* dyt, int lddt, mvd_stream_t stream);

/* distill_kl in ONE pass (dense C = 4 logits, NV % 4 == 0): loss_sum[0] (double, caller zeroes) += sum p_t (log p_t -
 * log p_s), and dys / dyt (bf16, either may be NULL) = gscale * dKL/dy with gscale = assumed upstream gradient * T^2 / numel
 * supplied by the caller.  mvd_rescale_bf16_pair multiplies both gradient tensors (n_elems bf16 each) by
 * gout[0] / assumed on the device and returns at once when that ratio is 1: exact autograd without a host sync. */
int mvd_kl_fused(const void* ys, const void* yt, long long NV, int C, float T, float gscale, double* loss_sum,
                 void* dys, void* dyt, mvd_stream_t stream);
int mvd_rescale_bf16_pair(void* a, void* b, long long n_elems, const float* gout, float assumed, mvd_stream_t stream);

/* ---- soft skeleton / clDice (training/loss/soft_skeleton.py:6-37) ------------------------------------------ */
/* fp32 volumes [B][D][H][W] */
int mvd_soft_erode(const float* in, float* out, int B, int D, int H, int W, mvd_stream_t stream);
int mvd_soft_dilate(const float* in, float* out, int B, int D, int H, int W, mvd_stream_t stream);
/* PyTorch tie rules: max_pool3d -> first maximum in scan order; torch.min -> 0.5/0.5 */
int mvd_soft_erode_bwd(const float* in, const float* gout, float* gin /*accumulated*/, int B, int D, int H, int W,
                       mvd_stream_t stream);
int mvd_soft_dilate_bwd(const float* in, const float* gout, float gscale, float* gin /*accumulated*/, int B, int D,
                        int H, int W, mvd_stream_t stream);
/* one skeleton level: delta = relu(E_j - dilate(E_j1)); skel_out = first ? delta : skel_in + relu(delta - skel_in*delta) */
int mvd_skel_update(const float* Ej, const float* Ej1, const float* skel_in, float* delta_out, float* skel_out,
                    int first, int B, int D, int H, int W, mvd_stream_t stream);
/* pointwise backward of the skeleton recursion for all levels: E/delta/skel are [L][N] stacks (L = iter+1),
 * g_skel [N] in, g_delta [L][N] out */
/* Fused forward pass over n_levels (1..4) consecutive soft_skel levels j0 .. j0+n-1, all iterations in shared memory:
 *   E_in = E_j0, skel_in = skeleton after level j0-1 (NULL when j0 == 0);
 *   E_next[l] (E_{j0+l+1}), delta[l], skel[l]: per-level outputs, each may be NULL except skel[n_levels-1].
 * Bit-identical to mvd_soft_erode + mvd_skel_update level by level.  (soft_skeleton.py:29-37) */
int mvd_soft_skel_fused(const float* E_in, const float* skel_in, int n_levels, float* const* E_next,
                        float* const* delta, float* const* skel, int B, int D, int H, int W, mvd_stream_t stream);
/* Fused backward of n_levels (1..2) consecutive soft_skel levels a .. a+n-1 with all scatter-adds in shared memory (no
 * global atomics) and delta recomputed from the E volumes:
 *   E[0..n]        : E_a ... E_{a+n} (E_0 = the input of soft_skel, E_{j+1} = erode(E_j): what mvd_soft_skel_fused stores)
 *   skel_prev[l]   : skeleton after level a+l-1; skel_prev[0] == NULL and first_is_level0 != 0 when a == 0
 *   G_in           : d loss / d skeleton entering level a+n-1 (the upstream gradient for the topmost launch)
 *   gE_top_in      : partial gradient of E_{a+n} written by the launch above (NULL for the topmost launch)
 *   gE_out, G_out  : gradient of E_a (the final input gradient when a == 0) and the chain state for the launch below */
int mvd_soft_skel_bwd_fused(const float* const* E, const float* const* skel_prev, int n_levels, int first_is_level0,
                            const float* G_in, const float* gE_top_in, float* gE_out, float* G_out, int B, int D, int H,
                            int W, mvd_stream_t stream);
int mvd_skel_chain_bwd(const float* delta, const float* skel, const float* g_skel, float* g_delta, int L,
                       long long N, mvd_stream_t stream);
/* gE_j += g_delta_j * [delta_j > 0];  gE_j1 (via dilate backward) -= same   (one level) */
int mvd_skel_level_bwd(const float* Ej1, const float* delta_j, const float* g_delta_j, float* gEj, float* gEj1,
                       int B, int D, int H, int W, mvd_stream_t stream);
/* sums[0..3] (double, caller zeroes) += sum(a*b), sum(a) , used for clDice's tprec / tsens */
int mvd_dot_sum(const float* a, const float* b, long long N, double* sums2, mvd_stream_t stream);
/* prob = softmax(logits)[:, channel] (fp32) and onehot = (target == channel) (fp32) */
int mvd_softmax_channel_fwd(const void* logits, int ld, const float* target, long long NV, int C, int channel,
                            float* prob, float* onehot, mvd_stream_t stream);
/* dlogits (bf16) = dprob * d softmax_channel / d logits */
int mvd_softmax_channel_bwd(const void* logits, int ld, const float* dprob, long long NV, int C, int channel,
                            void* dlogits, int ldd, mvd_stream_t stream);
/* clDice chain rule, device scalars only (no host sync): out4 from mvd_cldice_finalize, gout may be NULL (=1)
 *   seed:    g_skel[i] = gout * (out4[1]*y[i] + out4[2])
 *   combine: dprob[i]  = gE0[i] + gout * out4[3] * skel_y[i] */
int mvd_cldice_seed(const float* y, const float* out4, const float* gout, float* g_skel, long long N,
                    mvd_stream_t stream);
int mvd_cldice_combine(const float* gE0, const float* skel_y, const float* out4, const float* gout, float* dprob,
                       long long N, mvd_stream_t stream);
/* clDice scalar + chain-rule coefficients from the four sums: sums = [S(skel_p*y), S(skel_p), S(skel_y*p), S(skel_y)]
 * out[0] = loss; out[1] = dL/dS1, out[2] = dL/dS2, out[3] = dL/dS3   (fp32) */
int mvd_cldice_finalize(const double* sums4, float smooth, float* out4, mvd_stream_t stream);

/* ---- optimiser tail (MVDTrainer.py:978-979, 482-486) ------------------------------------------------------- */
/* multi-tensor tables live in device memory: ptrs = [n][3] (param, grad, momentum_buf) as uint64, numel = [n] int64,
 * chunk_tensor / chunk_offset = [n_chunks] (int32 tensor id, int64 element offset), chunk elements = 4096 */
int mvd_grad_sqnorm(const uint64_t* ptrs, const long long* numel, const int* chunk_tensor,
                    const long long* chunk_offset, int n_chunks, double* sqnorm /*caller zeroes*/,
                    mvd_stream_t stream);
/* clip coef = min(1, max_norm / (sqrt(sqnorm*gscale^2) + 1e-6)); g = grad*gscale*coef + wd*p; buf = mom*buf + g;
 * p -= lr * (g + mom*buf)   (torch.optim.SGD nesterov; gscale folds the DDP 1/world mean) */
int mvd_sgd_nesterov_clip(const uint64_t* ptrs, const long long* numel, const int* chunk_tensor,
                          const long long* chunk_offset, int n_chunks, const double* sqnorm, float gscale,
                          float max_norm, float lr, float weight_decay, float momentum, mvd_stream_t stream);

/* ---- utility ----------------------------------------------------------------------------------------------- */
/* out[c] = sum over NV voxels of g[v][c] (fp32; bias gradient of ConvTranspose3d) */
int mvd_channel_sum(const void* g, int ld, long long NV, int C, float* out, mvd_stream_t stream);
/* out[c] = sum_b stats[b][c0 + c][0] for c < n: turns the statistics a conv epilogue produced (mvd_conv3d_args.stats,
 * [B][C][2] doubles) into a per-channel sum, e.g. the bias gradient of a ConvTranspose3d from the sums of the gradient
 * tensor its consumer's dgrad just wrote (replaces a full pass of mvd_channel_sum over that tensor) */
int mvd_stats_channel_sum(const double* stats, int B, int C, int c0, int n, float* out, mvd_stream_t stream);
/* zero n regions of one fp32 buffer in a single launch: table (device) = n x (element offset, element count) int64.
 * The trainer clears the accumulate-type gradients of a step (conv biases in front of InstanceNorm, head weights) with it. */
int mvd_zero_regions(float* base, const long long* table_device, int n, mvd_stream_t stream);
/* cudaMemsetAsync(ptr, 0, bytes) on the stream: the per-step scratch pool */
int mvd_zero_bytes(void* ptr, size_t bytes, mvd_stream_t stream);
/* out[0] (+)= scale * in[0]: device-side double -> float scalar algebra (keeps the step free of host syncs) */
int mvd_scalar_axpy(const double* in, float scale, float* out, int accumulate, mvd_stream_t stream);
int mvd_add_bf16(void* dst, int ldd, const void* src, int lds, long long NV, int C, mvd_stream_t stream); /* dst += src */
#ifdef __cplusplus
}
#endif
#endif /* MVDSEG_H_ */
