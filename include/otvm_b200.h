/* otvm_b200 — C ABI of the B200 (sm_100a) kernels behind the OTVM per-frame inference hot path.
 *
 * The reference (Hongje/OTVM) is pure Python/PyTorch and defines no FFI; its "operator interface" is the
 * set of torch.nn / torch.nn.functional calls on the path (SURVEY.md §2.3, §8(b) last row).  Each entry
 * point below replaces the reference call sites it cites (paths relative to the reference root).
 *
 * Conventions
 *   - plain C: raw device pointers + sizes, no torch types.  The caller owns every buffer (inputs,
 *     outputs, workspaces); the library never allocates, frees or retains device memory.
 *   - activations are NHWC ("pixel-major"): element (n,y,x,c) of a tensor lives at
 *     base[((n*H + y)*W + x) * ld + c]; `ld` >= C lets a producer write straight into a channel slice of a
 *     wider concat buffer (torch.cat is never materialised by a copy).
 *   - dtype: OTVM_F32 (strict fp32 arithmetic, FFMA), OTVM_BF16 (bf16 storage, tcgen05 tensor cores with fp32
 *     accumulation) or a SPLIT format OTVM_SPLIT_DTYPE(planes, plane_stride_bytes), planes = 2 or 3: every
 *     activation is the sum of `planes` bf16 tensors ("planes": bf16(v), bf16(v - plane0), ...; 16 / 24 significant
 *     bits) that are identical views `plane_stride_bytes` apart, i.e. plane k of the tensor at address p lives at
 *     p + k * plane_stride_bytes.  All split tensors passed to one call share that stride (allocate them from one
 *     [planes][bytes] arena; the stride is a multiple of 4096).  The tensor cores multiply the planes pairwise
 *     (3 products for 2 planes, 6 for 3: fp32-grade results at tensor-core rate); this is what lets the tcgen05 path
 *     meet the reference's fp32 outputs to 1e-2 (2 planes) / 1e-3 (3 planes), which plain bf16 storage cannot
 *     (DESIGN.md section 4).  Statistics, softmax state and biases are always fp32.
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), re-entrant per stream, and
 *     returns 0 on success or a negative code (see otvm_strerror); it never throws.
 */
#ifndef OTVM_B200_H
#define OTVM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OTVM_ABI_VERSION 5
#if defined(__GNUC__)
#define OTVM_API __attribute__((visibility("default")))
#else
#define OTVM_API
#endif

enum { OTVM_F32 = 0, OTVM_BF16 = 1, OTVM_BF16X2 = 2, OTVM_BF16X3 = 3 };
/* dtype word of a split format: low byte = OTVM_BF16X2 / X3, upper bits = plane stride in 4096-byte units */
#define OTVM_SPLIT_DTYPE(planes, plane_stride_bytes) \
  ((int32_t)((planes) == 3 ? OTVM_BF16X3 : OTVM_BF16X2) | (int32_t)(((int64_t)(plane_stride_bytes) / 4096) << 8))
enum { OTVM_ACT_NONE = 0, OTVM_ACT_RELU = 1, OTVM_ACT_LEAKY = 2 };     /* LeakyReLU slope 0.01 (nn default) */
enum {
  OTVM_OK = 0, OTVM_ERR_ARG = -1, OTVM_ERR_CUDA = -2, OTVM_ERR_UNSUPPORTED = -3, OTVM_ERR_WORKSPACE = -4
};

OTVM_API int otvm_version(void);
OTVM_API const char* otvm_strerror(int code);
/* last CUDA error string seen by this library on the calling thread (for OTVM_ERR_CUDA) */
OTVM_API const char* otvm_last_cuda_error(void);
/* number of CUDA kernels this library has launched since it was loaded (all threads) */
OTVM_API int64_t otvm_launch_count(void);
/* 1 when the device is compute capability 10.x (tcgen05 / TMA paths usable) */
OTVM_API int otvm_device_is_sm100(int device);
/* Programmatic dependent launch between consecutive kernels of a stream (default on; env OTVM_PDL=0 disables):
 * the next kernel's prologue overlaps the tail of the previous one; also recorded as programmatic edges when
 * the stream is being captured into a CUDA graph.
 * CONTRACT: parameters (convolution weight / bias, GroupNorm gamma / beta) are treated as constants of the stream:
 * kernels read them BEFORE they wait for their predecessor (weight tiles are prefetched during the previous
 * kernel's tail).  A caller that rewrites a parameter buffer on the device must synchronise the stream (or call
 * otvm_set_pdl(0)) between that write and the next call that reads it.  Activations carry no such restriction. */
OTVM_API void otvm_set_pdl(int enabled);
/* Sticky device-side error flags of the current device (synchronises the device; `clear` != 0 resets them).
 *   bit 0: the grid-wide barrier of a GroupNorm-fused convolution (otvm_conv_params.gn_gamma) timed out because a CTA
 *          of its grid never became resident (something else held the SMs); that launch's output is invalid, the
 *          stream and the process remain usable.  Returns -1 when the flags cannot be read. */
OTVM_API int otvm_device_error_flags(int clear);
/* cudaMemsetAsync(ptr, 0, bytes) on `stream` (e.g. the per-frame GroupNorm statistics arena) */
OTVM_API int otvm_zero_async(void* ptr, int64_t bytes, void* stream);

/* ---- convolution ------------------------------------------------------------------------------------
 * Replaces every nn.Conv2d / F.conv2d on the path: models/trimap/STM.py:17-19,37-43,107,122-127,169-170;
 * models/alpha/FBA/layers_WS.py:22 (weights arrive already standardised, layers_WS.py:15-21 is folded at
 * load time because weights are constant in eval); models/alpha/FBA/models.py:303-347,399-415.
 * Eval-mode BatchNorm (STM.py:44,80 and torchvision Bottleneck) is folded into weight/bias by the host.
 *
 *   out[p, co] = act( sum_{ky,kx,ci} w[co,ky,kx,ci] * pre(in[p*stride - pad + (ky,kx)*dil, ci]) + bias[co]
 *                     + res[p, co] )                  pre = ReLU when relu_in (STM ResBlock, STM.py:24-25)
 */
typedef struct {
  const void* in;      int64_t in_ld;             /* NHWC input, channel stride per pixel (elements)        */
  int32_t N, H, W, Cin;
  const void* weight;                             /* [Cout][KH][KW][Cin], same element format as activations */
                                                  /* (split formats: planes `w_plane_stride` elements apart) */
  const float* bias;                              /* [Cout] fp32 or NULL                                    */
  int32_t Cout, KH, KW, stride, pad, dil;
  void* out;           int64_t out_ps, out_cs;    /* out element (p,co) at out[p*out_ps + co*out_cs]        */
  const void* res;     int64_t res_ld;            /* optional residual (NHWC, added before act) or NULL     */
  void* out_relu;      int64_t out_relu_ld;       /* optional second NHWC output = ReLU(out) or NULL        */
  int32_t act;                                    /* OTVM_ACT_*                                             */
  int32_t relu_in;                                /* apply ReLU to the input while loading                  */
  int32_t dtype;                                  /* element format of in, weight, out, res, out_relu        */
  int32_t out_f32;                                /* 1: `out` is fp32 even when dtype is bf16 (heads)       */
  double* gn_stats;                               /* optional [N][32][2] (sum, sumsq) accumulated over out  */
  void* workspace;     int64_t workspace_bytes;   /* optional fp32 scratch: enables split-K for small grids */
  int32_t gn_stats_zeroed;                        /* 1: caller already zeroed gn_stats on this stream (one   */
  float gn_eps;                                   /*    otvm_zero_async over an arena instead of a memset    */
                                                  /*    node in front of every convolution)                  */
  /* Fused GroupNorm(32, Cout) (optional): with gn_gamma/gn_beta set the kernel normalises its own output,
   * out = act(GN(conv + bias) * gamma + beta + res), replacing the otvm_gn_apply pass: statistics in the epilogue,
   * a grid-wide barrier, then normalise from the accumulators still in tensor memory.  Requires a tcgen05 shape
   * whose grid is one co-resident wave, gn_stats = a zeroed slot of >= 65 doubles ([32][2] sums + barrier counter,
   * gn_stats_zeroed = 1) and no out_relu; ask otvm_conv2d_can_fuse_gn first (otvm_conv2d returns
   * OTVM_ERR_UNSUPPORTED otherwise and launches nothing). */
  const float* gn_gamma; const float* gn_beta;
  int64_t w_plane_stride;                         /* split formats: elements between the planes of `weight`  */
  /* Channels per GroupNorm group; 0 = Cout / 32 (GroupNorm(32, Cout) over this call's channels).  A caller can run
   * a wide layer as channel SLICES -- one call per slice with the slice's weights / gamma / beta / out / res views,
   * gn_group_ch = full Cout / 32 and its own statistics slot: the slice then holds Cout / gn_group_ch whole groups
   * (statistics [Cout / gn_group_ch][2]).  Groups never straddle slices, so the result is that of the unsliced
   * layer, and a slice whose grid is one co-resident wave can take the fused path (gn_gamma) where the whole layer
   * could not (FBA layer4: 512 -> 2048 at 1/8 resolution).  tcgen05 path only. */
  int32_t gn_group_ch;
  /* Weight groups; 0 / 1 = one filter bank.  groups = G > 1: `weight` holds G banks [G][Cout][KH][KW][Cin] (per plane)
   * and `bias` [G][Cout]; image n of the batch (N == G) is convolved with bank n.  One launch then runs G layers of
   * identical shape on G inputs -- the two STM encoders (Encoder_Q on the current frame, Encoder_M on the previous
   * one, STM.py:33-102: same ResNet-50 stages, different weights) share every launch instead of competing on two
   * streams.  tcgen05 path only; no GroupNorm, no split-K. */
  int32_t groups;
} otvm_conv_params;
OTVM_API int otvm_conv2d(const otvm_conv_params* p, void* stream);
/* 1 when otvm_conv2d can run this problem with the GroupNorm fused into the convolution kernel (see gn_gamma) */
OTVM_API int otvm_conv2d_can_fuse_gn(const otvm_conv_params* p);
/* 1 when otvm_conv2d would run this problem on the tcgen05 implicit-GEMM kernel (else the FFMA kernel) */
OTVM_API int otvm_conv2d_uses_tensor_cores(const otvm_conv_params* p);

/* ---- GroupNorm(32, C) -------------------------------------------------------------------------------
 * Replaces nn.GroupNorm(32,C) (models/alpha/FBA/layers_WS.py:26-27; FBA/models.py:272-276) together with
 * the ReLU/LeakyReLU and residual add that follow it (resnet_GN_WS.py:32-48,69-88; FBA/models.py:305-336).
 * stats: [N][32][2] doubles (sum, sum of squares), zeroed by otvm_gn_stats itself before accumulating.
 *   y = act( (x - mean_g) * rstd_g * gamma[c] + beta[c] + res )           eps = 1e-5
 */
OTVM_API int otvm_gn_stats(const void* x, int64_t ld, int32_t N, int32_t HW, int32_t C, int32_t dtype,
                  double* stats, void* stream);
OTVM_API int otvm_gn_apply(const void* x, int64_t ld, int32_t N, int32_t HW, int32_t C, int32_t dtype,
                  const double* stats, const float* gamma, const float* beta, float eps,
                  const void* res, int64_t res_ld, int32_t act, void* out, int64_t out_ld, void* stream);

/* ---- resampling / pooling ---------------------------------------------------------------------------
 * F.interpolate(mode='bilinear', align_corners=False): STM.py:115,136; FBA/models.py:358-361,366,371,376.
 * out[.., c] = (add ? add[.., c] : 0) + bilinear(in)[.., c]; out_relu (optional) = ReLU(out), the input of the
 * ResBlock that follows (STM.py:24,116); output may be fp32 NCHW planes (out_nchw_f32) for the STM logits. */
OTVM_API int otvm_upsample_bilinear(const void* in, int64_t in_ld, int32_t N, int32_t Hi, int32_t Wi, int32_t C,
                           int32_t Ho, int32_t Wo, const void* add, int64_t add_ld,
                           void* out, int64_t out_ld, void* out_relu, int64_t out_relu_ld,
                           int32_t dtype, int32_t out_nchw_f32, void* stream);
/* nn.MaxPool2d(3, 2, 1): STM.py:47,83; resnet_GN_WS.py:98 (indices are never used, FBA/models.py:338) */
OTVM_API int otvm_maxpool3x3s2(const void* in, int64_t in_ld, int32_t N, int32_t H, int32_t W, int32_t C,
                      void* out, int64_t out_ld, int32_t dtype, void* stream);
/* nn.AdaptiveAvgPool2d(s), s in {1,2,3,6} in ONE pass over the input (FBA/models.py:302):
 * out is [N][50][C] (1 + 4 + 9 + 36 cells, scale-major); scratch: N*H*12*C floats (per-row column-bin sums) */
OTVM_API int otvm_ppm_pool(const void* in, int64_t in_ld, int32_t N, int32_t H, int32_t W, int32_t C,
                  void* out, float* scratch, int32_t dtype, void* stream);

/* ---- STM space-time memory read ---------------------------------------------------------------------
 * Replaces Memory.forward, models/trimap/STM.py:144-163 (bmm :153, /sqrt(De) :154, softmax over THW :155,
 * bmm :158; the torch.cat at :161 disappears because `out` is channels [0,Do) of the decoder input buffer
 * whose channels [Do,2Do) the KV_Q value conv writes).  ONE fused kernel with online softmax per split of
 * the memory axis + a log-sum-exp combine; the [THW x HW] affinity is never materialised.
 *   keys  : [M][De]   pixel-major rows, M = T*h*w memory locations (ld = De)
 *   vals  : [Do][ldv] channel-major, location m of channel c at vals[c*ldv + m]
 *   query : [HW][q_ld] (first De channels used);   out : [HW][out_ld] (first Do channels written)
 * workspace: fp32, at least otvm_memory_read_workspace(...) bytes.
 */
typedef struct {
  const void* keys; const void* vals; int64_t ldv;
  const void* query; int64_t q_ld;
  void* out; int64_t out_ld;
  int32_t M, HW, De, Do;
  int32_t dtype;
  void* workspace; int64_t workspace_bytes;
  int32_t force_simt;                           /* 1: use the fp32 FFMA kernel even for bf16 inputs (tests) */
  float* lse;                                   /* optional [HW]: log2-sum-exp of the scaled logits per query    */
                                                /* (saved for otvm_memory_read_backward), or NULL              */
} otvm_read_params;
OTVM_API int64_t otvm_memory_read_workspace(int32_t M, int32_t HW, int32_t De, int32_t Do, int32_t dtype);
OTVM_API int otvm_memory_read(const otvm_read_params* p, void* stream);

/* Backward of the read (the stage-4 training step differentiates through
 * Memory.forward, train.py:349-375).  Recomputes 32 x 32 tiles of the affinity from the operands and the saved
 * log-sum-exp (nothing of size THW x HW is stored), fp32 arithmetic.  Operands as in the forward (dtype OTVM_F32 or
 * OTVM_BF16); out / dout: forward output and its gradient, fp32 [HW][ld]; gradients fp32:
 *   dkeys [M][De], dvals [Do][dldv], dquery [HW][De] (all overwritten). */
typedef struct {
  const void* keys; const void* vals; int64_t ldv;
  const void* query; int64_t q_ld;
  const float* out; int64_t out_ld;
  const float* dout; int64_t dout_ld;
  const float* lse;
  float* dkeys; float* dvals; int64_t dldv; float* dquery;
  int32_t M, HW, De, Do;
  int32_t dtype;
} otvm_read_bwd_params;
OTVM_API int otvm_memory_read_backward(const otvm_read_bwd_params* p, void* stream);

/* ---- frame glue (EvalModel.forward, models/alpha/model.py:391-512) ----------------------------------
 * preprocess_gt + make_trimap_gt (models/alpha/model.py:342-362,380-389): BGR->RGB flip, 1/255, composite,
 * unknown mask, (2r+1)^2 max-pool dilation, one-hot, centred pad to a multiple of 32 (:408-410).
 *   a [H*W], fg/bg [3][H*W] fp32 planar (the eval.py tensors);  img: [Hp*Wp][4] fp32 RGB0 in [0,1];
 *   scaled_img: [3][H*W] fp32 planar un-padded (first return value of EvalModel.forward);
 *   tri3: [Hp*Wp][4] fp32 one-hot (bg, unknown, fg, 0), padding = bg;
 *   imgn: [Hp*Wp][imgn_ld] dtype, (img - mean) / std + one zero channel: the STM query-encoder input
 *   (STM.py:93); mean_std: HOST pointer, 3 means then 3 stds;  scratch: 2*H*W bytes. */
OTVM_API int otvm_preprocess(const float* a, const float* fg, const float* bg, int32_t H, int32_t W,
                    int32_t Hp, int32_t Wp, int32_t pad_top, int32_t pad_left, int32_t radius,
                    const float* mean_std, float* img, float* scaled_img, float* tri3,
                    void* imgn, int64_t imgn_ld, int32_t dtype, uint8_t* scratch, void* stream);

/* make_trimap + trimap_transform (models/alpha/model.py:40-53, utils/utils.py:12-39) and the 11-channel FBA
 * input (models/alpha/model.py:414,445): optional softmax over the 3 logits, argmax classes, EXACT Euclidean
 * distance transform of the bg and fg masks on the device (replaces cv2.distanceTransform on the host),
 * three Gaussians per mask, soft bg/fg channels, ImageNet normalisation of the image.
 *   tri_in : [Hp*Wp][tri_ld] fp32, 3 logits (is_logit=1) or 3 probabilities per pixel
 *   img    : [Hp*Wp][4] fp32 RGB0 in [0,1] (otvm_preprocess);  mean_std: HOST pointer, 3 means then 3 stds
 *   x11    : [Hp*Wp][x11_ld>=16] dtype — 3 normalised RGB + 6 distance channels + soft bg + soft fg + 5 zeros
 *   cat_dst: optional [Hp*Wp][cat_ld] dtype, 8 channels written per pixel: normalised RGB, RGB, soft bg, soft fg
 *            (channels 64..71 of the conv_up4 / refine input concat, FBA/models.py:377-378,418)
 *   extras : [Hp*Wp][8] fp32 — RGB in [0,1] (3), soft bg, soft fg, 3 class probabilities (bg, un, fg)
 *   d2     : [2][Hp*Wp] int32 squared distances to the nearest bg / fg pixel (exposed for bit-exact tests)
 *   scratch: 2*Hp*Wp int32;  seeds: 2*Hp*Wp bytes */
OTVM_API int otvm_trimap_encode(const float* tri_in, int64_t tri_ld, int32_t is_logit, const float* img,
                       int32_t Hp, int32_t Wp, const float* mean_std, void* x11, int64_t x11_ld,
                       void* cat_dst, int64_t cat_ld, int32_t dtype,
                       float* extras, int32_t* d2, int32_t* scratch, uint8_t* seeds, void* stream);

/* Exact squared Euclidean distance transform on its own (utils/utils.py:21, cv2.distanceTransform with
 * DIST_L2 / DIST_MASK_PRECISE before its sqrt): d2[p] = min over seed pixels q (seed[q] != 0) |p-q|^2,
 * INT32_MAX/2 when there is no seed. */
OTVM_API int otvm_edt_sq(const uint8_t* seed, int32_t H, int32_t W, int32_t* d2, int32_t* scratch, void* stream);

/* clamp / sigmoid / fba_fusion (models/alpha/FBA/models.py:279-288,383-390,425-431).
 *   raw  : [P][raw_ld] dtype, channels 0..6 = (alpha, F rgb, B rgb) pre-activation, 7..9 = trimap logits
 *   img  : [P][8] fp32 extras (RGB first)
 *   out7 : [P][8] fp32 fused (alpha, F, B, 0);  alpha_dst: optional dtype buffer, alpha written at
 *   alpha_dst[p*alpha_ld] (channel slot of the refine input concat, FBA/models.py:418) */
OTVM_API int otvm_fba_head(const void* raw, int64_t raw_ld, int32_t dtype, int32_t raw_f32, const float* extras,
                  int64_t P, float* out7, void* alpha_dst, int64_t alpha_ld, void* stream);

/* The 1x1 head convolution on a 16-channel activation + otvm_fba_head in ONE pointwise pass: replaces
 * nn.Conv2d(16, 7, 1) `conv_up4.4` (FBA/models.py:347) / nn.Conv2d(16, 10, 1) `pred.4` (:415) and the fusion behind them.
 *   x   : [P][x_ld] dtype, 16 channels;   w: [Cout][16] fp32, bias: [Cout] fp32 or NULL, 7 <= Cout <= 12
 *   raw : [P][raw_ld] fp32 (raw_ld = 8 or 12): raw[p][co] = bias[co] + sum_c w[co][c] x[p][c], columns >= Cout zero
 *   extras / out7 / alpha_dst / alpha_ld as in otvm_fba_head (the fusion reads raw[p][0..6]) */
OTVM_API int otvm_head_conv_fba(const void* x, int64_t x_ld, int32_t dtype, const float* w, const float* bias, int32_t Cout,
                       float* raw, int64_t raw_ld, const float* extras, int64_t P, float* out7, void* alpha_dst,
                       int64_t alpha_ld, void* stream);

/* softmax of the refined trimap logits (models/alpha/model.py:460), the 20-channel memorize input
 * cat(tri3, alpha, hid16) + frame (models/trimap/model.py:231, STM.py:56-67: 22 = 3 normalised RGB + unknown
 * + fg + alpha + 16 hidden, padded to mem_ld), and the cropped planar outputs eval.py reads (:495-508).
 *   raw10 : [P][raw_ld] fp32 refine head (7 fused inputs ignored here, 7..9 trimap logits)
 *   fused : [P][8] fp32 (alpha first);  hid: [P][hid_ld] dtype 16 channels;  extras: [P][8] fp32
 *   mem_in: [P][mem_ld>=24] dtype (22 channels + zeros), NULL on the last frame;  mean_std: HOST pointer;
 *    alpha_out [H*W] fp32;  trimap_out [3][H*W] fp32 (cropped, planar) */
OTVM_API int otvm_frame_outputs(const float* raw10, int64_t raw_ld, const float* fused, const void* hid,
                       int64_t hid_ld, const float* extras, int32_t Hp, int32_t Wp, int32_t H, int32_t W,
                       int32_t pad_top, int32_t pad_left, const float* mean_std, void* mem_in, int64_t mem_ld,
                       int32_t dtype, float* alpha_out, float* trimap_out, void* stream);

/* ---- frame I/O around the loop -----------------------------------------
 * The reference decodes each frame on the host into fp32 tensors (dataset.py:857-920: cv2.imread, np.float32, / 255,
 * HWC -> CHW) and reads the matte back as fp32 before converting it (eval.py:209: (alphas * 255).byte()).  Here the
 * decoded 8-bit images are uploaded as they are and unpacked on the device, and the matte returns as 8 bits.
 *   fg  : [H*W][fg_channels] u8, BGR(A) as cv2.imread(IMREAD_UNCHANGED) returns it; bg: [H*W][3] u8 BGR
 *   a   : [H*W] fp32 = A / 255 (1.0 without an alpha channel);  fg_out / bg_out: [3][H*W] fp32 planar BGR in 0..255
 *         -- exactly the tensors EvalModel.forward takes (eval.py:162-175), bit-identical to the host decode */
OTVM_API int otvm_unpack_frame_u8(const uint8_t* fg, int32_t fg_channels, const uint8_t* bg, int32_t H, int32_t W,
                         float* a, float* fg_out, float* bg_out, void* stream);
/* out[p] = (uint8)(alpha[p] * 255): the PNG eval.py:217 writes, bit-identical to (alphas * 255).byte() */
OTVM_API int otvm_alpha_to_u8(const float* alpha, int64_t P, uint8_t* out, void* stream);

/* dtype conversion / layout helpers used at the boundary (NCHW fp32 <-> NHWC dtype) */
OTVM_API int otvm_nchw_to_nhwc(const float* in, int32_t N, int32_t C, int32_t HW, void* out, int64_t out_ld,
                      int32_t dtype, void* stream);
OTVM_API int otvm_nhwc_to_nchw(const void* in, int64_t in_ld, int32_t N, int32_t C, int32_t HW, float* out,
                      int32_t dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OTVM_B200_H */
