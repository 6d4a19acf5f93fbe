/*
 * ast.h - C ABI of libast_b200.so: the B200 (sm_100a) kernels behind the perceptual-loss
 * training step of edogariu/artist-style-transfer.
 *
 * Drop-in boundary: the reference has no native code; the "FFI" it binds for this
 * path is the set of PyTorch library calls listed next to each entry point below.  Host code stays
 * Python/PyTorch (artist_style_transfer_b200/{cnn,train_cnn}.py mirror the reference modules) and
 * reaches these functions through ctypes with raw device pointers - no torch types cross the ABI.
 *
 * Conventions
 *   - every function returns 0 on success, <0 for an argument/shape/unsupported-config error
 *     (never a silent fallback), >0 for a cudaError_t / CUresult; ast_last_error() has the text.
 *   - the caller owns every buffer (inputs, outputs, workspaces) and keeps it alive until the work
 *     queued on `stream` has completed; the library never allocates device memory and keeps no
 *     pointers after return.  Launches are asynchronous on `stream`.
 *   - images are N x H x W x C views with explicit element strides (ast_image); the internal layout
 *     is NHWC (sc == 1), but NCHW tensors of the reference API are described with sc == H*W.
 */
#ifndef AST_B200_H
#define AST_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AST_ABI_VERSION 4

typedef enum { AST_F32 = 0, AST_BF16 = 1, AST_TF32 = 2 /* fp32 storage rounded to TF32; pack_weights only */ } ast_dtype;

/* strided 4-D view; strides in ELEMENTS.  */
typedef struct {
  void*   ptr;
  int32_t dtype;          /* ast_dtype */
  int32_t n, h, w, c;
  int64_t sn, sh, sw, sc;
} ast_image;

#define AST_MAX_TAPS 81

/* A "gather" convolution: for every point (i,j) of an Mi x Mj iteration grid
 *   out[n, oy0 + so*i, ox0 + so*j, co] = epi( sum_t sum_ci  in[n, si*i + dy[t], si*j + dx[t], ci] * W[t][co][ci] )
 * Out-of-range input coordinates read 0 (or are mirrored when AST_CONV_REFLECT is set).
 * This single primitive expresses (tap tables are built by the host mirror, conv_geometry.py):
 *   nn.Conv2d stride 1/2 after nn.ReflectionPad2d      cnn.py:58,63,73-74   (si = stride)
 *   its data gradient (aten::convolution_backward)      train_cnn.py:333     (so = stride, 4 phases)
 *   nn.ConvTranspose2d k3 s2 p1 op1 and k1 s1           cnn.py:107-109,119   (so = 2, 4 sub-pixel phases)
 *   torchvision VGG16 Conv2d(3x3, pad 1) + ReLU         train_cnn.py:54,72-73
 *   the Gram backward  dF = (dG + dG^T) F / (CHW)       train_cnn.py:107,333 (1 tap, per-image weights)
 */
typedef struct {
  int32_t mi, mj;          /* iteration grid per image */
  int32_t si, so;          /* input / output coordinate multipliers */
  int32_t oy0, ox0;        /* output phase offset */
  int32_t ntaps;
  int32_t flags;           /* AST_CONV_* */
  int16_t dy[AST_MAX_TAPS];
  int16_t dx[AST_MAX_TAPS];
  int64_t w_img_stride;    /* elements between per-image weight sets, 0 = shared weights */
  float*  stats;           /* optional fp32 [n][cout][2] (sum x, sum x^2 of the fp32 results, BEFORE bias/activation),
                              accumulated by the tensor-core epilogue: InstanceNorm statistics fused into the producing
                              conv (cnn.py:63,68).  Zeroed by the caller; NULL = off; tensor-core launches only. */
  const ast_image* pooled; /* optional [n, mi/2, mj/2, cout]: nn.MaxPool2d(2, 2) of the epilogue result, written by the same
                              kernel (torchvision VGG16 features idx 4/9/16 after the conv+ReLU before them,
                              train_cnn.py:54,72-73).  Plain stride-1 launches (so = 1, no phase offset) of the
                              weight-stationary tensor-core kernel only: any other request is an error.  With
                              AST_CONV_POOL_ONLY the full-resolution `out` is not written (no-grad content branch). */
} ast_gather_geom;

#define AST_CONV_RELU     1   /* epilogue max(v,0)                       (nn.ReLU, train_cnn.py VGG idx 1,3,...)   */
#define AST_CONV_REFLECT  2   /* mirror out-of-range input coordinates   (nn.ReflectionPad2d, cnn.py:58)          */
#define AST_CONV_POOL_ONLY 16  /* with geom.pooled: skip the store of the full-resolution output                      */
#define AST_CONV_TENSOR   4   /* request the tcgen05/TMA kernel; error if the shape is not supported              */
#define AST_CONV_ROUND_TF32 8 /* round the fp32 result to TF32 (cvt.rna) so a kind::tf32 consumer sees exact operands */

/* weights: packed [ntaps][cout][cin], same dtype as `in`.  bias: fp32[cout] or NULL.
 * in_shift: fp32[cin] added to in-range inputs before the product (VGG mean shift, train_cnn.py:300-301) or NULL.
 * add:  optional image added before the activation (tap-gradient accumulation in the VGG backward).
 * mask: optional image; out = (mask > 0) ? v : 0 after the activation (ReLU backward, threshold_backward). */
int ast_conv_gather(const ast_image* in, const void* weights, const float* bias, const float* in_shift,
                    const ast_image* add, const ast_image* mask, const ast_image* out,
                    const ast_gather_geom* geom, void* stream);

/* Weight gradient of the same gather convolution (aten::convolution_backward, filter part):
 *   dw[tap_off[t] + co*s_co + ci*s_ci] += sum_{n,i,j} gout[n, oy0+so*i, ox0+so*j, co] * x[n, si*i+dy[t], si*j+dx[t], ci]
 * dw is fp32 and must be zeroed (or hold a running sum) by the caller; accumulation uses fp32 atomics. */
int ast_wgrad_gather(const ast_image* x, const ast_image* gout, float* dw, const int32_t* tap_off,
                     int64_t s_co, int64_t s_ci, const ast_gather_geom* geom, void* stream);

/* Re-pack fp32 master weights into the [ntaps][a][b] operand layout of ast_conv_gather:
 *   dst[t][ia][ib] = src[tap_off[t] + ia*s_a + ib*s_b]      (dst dtype = ast_dtype) */
int ast_pack_weights(const float* src, const int32_t* tap_off, int32_t ntaps, int32_t a, int32_t b,
                     int64_t s_a, int64_t s_b, void* dst, int32_t dst_dtype, void* stream);

/* Same with a two-level inner index and zero padding (layouts of the 3-channel layers, see ast_row_im2col):
 *   dst[t][ia][ib] = (ia < a_valid && ib < b_valid) ? src[tap_off[t] + ia*s_a + (ib / b0)*s_b1 + (ib % b0)*s_b0] : 0 */
int ast_pack_weights_ex(const float* src, const int32_t* tap_off, int32_t ntaps, int32_t a, int32_t a_valid,
                        int32_t b, int32_t b_valid, int32_t b0, int64_t s_a, int64_t s_b1, int64_t s_b0,
                        void* dst, int32_t dst_dtype, void* stream);

/* "Row im2col" of a thin (3-channel) image so that its k x k convolution becomes a k-tap vertical convolution over
 * kw*C (zero-padded to out.c) channels that TMA / tcgen05 can consume:
 *   out[n, y, x, d*C + c] = src[n, Y(y - py), X(x + sign*d - px), c] (+ shift[c]),  d in [0, kw)
 * Out-of-range source coordinates are mirrored (reflect != 0: nn.ReflectionPad2d, cnn.py:58) or read as 0 (zero padding
 * of VGG conv1_1 applied AFTER the mean shift, train_cnn.py:300-301).  round_tf32 rounds fp32 outputs to TF32. */
int ast_row_im2col(const ast_image* src, const ast_image* out, const float* shift, int32_t kw, int32_t sign,
                   int32_t px, int32_t py, int32_t reflect, int32_t round_tf32, void* stream);

/* out[n, y, x, d*C + j] = src[n, y + sign*d - py, x, j] (0 outside), d in [0, kh): folds the kh vertical taps of an NHWC
 * tensor into channels so that a kh-tap ast_wgrad_gather becomes one tap over kh*C channels (thin 9x9 layers). */
int ast_unfold_rows(const ast_image* src, const ast_image* out, int32_t kh, int32_t sign, int32_t py, void* stream);

/* Finishes a thin-OUTPUT k x k convolution (cnn.py:39, the 32->3 9x9 layer) whose kw horizontal taps were computed as
 * kw*C output channels of a k-tap vertical ast_conv_gather:
 *   out[n, y, x, c] = bias[c] + sum_{d < kw} part[n, y, x + d, d*C + c]      (optional ReLU)
 * part: fp32 NHWC [n, h, w + kw - 1, >= kw*C] with dense rows; out: any dtype / strides (e.g. an NCHW view). */
int ast_fold_rows(const ast_image* part, const ast_image* out, const float* bias, int32_t kw, int32_t relu, void* stream);

/* nn.InstanceNorm2d(affine=True), eps 1e-5, biased variance (cnn.py:68,114).
 * stats: mean[n*c], rstd[n*c] (fp32).  workspace: ast_instnorm_workspace_bytes(). */
int64_t ast_instnorm_workspace_bytes(int32_t n, int32_t c);
int ast_instnorm_stats(const ast_image* x, float* mean, float* rstd, float eps, void* workspace, void* stream);
/* mean/rstd from the (sum x, sum x^2) pairs a conv epilogue accumulated (ast_gather_geom.stats): [n*c][2] -> [n*c]. */
int ast_instnorm_finalize(const float* sums, int32_t n, int32_t c, int32_t hw, float eps, float* mean, float* rstd,
                          void* stream);
/* y = gamma*(x-mean)*rstd + beta (+ residual) (ReLU if relu), written to the interior of `out`, which is an
 * (h+2*pad) x (w+2*pad) image whose border is filled by mirroring (the next layer's nn.ReflectionPad2d). */
int ast_instnorm_apply(const ast_image* x, const float* mean, const float* rstd, const float* gamma,
                       const float* beta, const ast_image* residual, const ast_image* out, int32_t pad,
                       int32_t relu, void* stream);
/* Backward of (ReflectionPad o ReLU o (+residual) o InstanceNorm).
 *   g'   = (fold_reflect(gpad, pad) + gextra) * (relu ? y>0 : 1)
 *   s1[n,c] = sum g' ; s2[n,c] = sum g'*xhat                       (pass 1, ast_instnorm_bwd_stats)
 *   dx   = gamma*rstd*(g' - s1/HW - xhat*s2/HW) ; gtotal = g'      (pass 2, ast_instnorm_bwd_apply)
 * gpad may be NULL (no padded consumer), gextra may be NULL, gtotal may be NULL. */
int ast_instnorm_bwd_stats(const ast_image* x, const float* mean, const float* rstd, const float* gamma,
                           const float* beta, const ast_image* gpad, int32_t pad, const ast_image* gextra,
                           int32_t relu, float* s1, float* s2, void* stream);
int ast_instnorm_bwd_apply(const ast_image* x, const float* mean, const float* rstd, const float* gamma,
                           const float* beta, const ast_image* gpad, int32_t pad, const ast_image* gextra,
                           int32_t relu, const float* s1, const float* s2, const ast_image* dx,
                           const ast_image* gtotal, void* stream);

/* nn.MaxPool2d(2,2) of torchvision vgg16.features idx 4/9/16 and its backward fused with the tap-gradient add
 * and the ReLU mask of the producing layer:  gx = (route(gy) + gadd) * (x > 0). */
int ast_maxpool2_fwd(const ast_image* x, const ast_image* y, void* stream);
int ast_maxpool2_bwd(const ast_image* x, const ast_image* y, const ast_image* gy, const ast_image* gadd,
                     const ast_image* gx, void* stream);

/* gram(), train_cnn.py:103-107:  G[n] = F[n] F[n]^T * scale, F = x viewed as (c, h*w).  G fp32 [n][c][c], zeroed here. */
int ast_gram(const ast_image* x, float* g, float scale, int32_t flags, void* stream);

/* nn.MSELoss pieces (train_cnn.py:249,307,323): loss[0] += scale * sum((a-b)^2) ; grad = gscale*(a-b) (optional).
 * a/b/grad are images of identical logical shape. */
int ast_mse(const ast_image* a, const ast_image* b, float* loss, float scale, const ast_image* grad,
            float gscale, void* stream);

/* dst = src converted/re-strided (NCHW fp32 <-> NHWC bf16/fp32), optional per-channel shift, reflect pad. */
int ast_copy_image(const ast_image* src, const ast_image* dst, const float* shift, int32_t pad, void* stream);
/* acc(fp32 image) += x   ('smartaverage' in-place feature sum, train_cnn.py:239) */
int ast_accumulate(const ast_image* x, const ast_image* acc, void* stream);

/* out = (a + b) * (mask > 0)   (b may be NULL; nn.ReLU backward on an incoming tap gradient) */
int ast_mask_add(const ast_image* a, const ast_image* b, const ast_image* mask, const ast_image* out, void* stream);

/* bitmask of the tensor-core kernels compiled into this build: 1 = tcgen05 gather conv, 2 = tcgen05 Gram/wgrad */
int ast_capabilities(void);
const char* ast_last_error(void);
int ast_abi_version(void);
/* number of kernel launches issued through this library by the calling process (bench.py gpu_launches) */
int64_t ast_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
