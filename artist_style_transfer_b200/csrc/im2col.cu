// Helpers that let the 3-channel ends of both networks run on the tcgen05 kernels:
//
//  * ast_row_im2col: "row im2col" of a thin image.  out[n, y, x, d*C + c] = src[n, Y(y), X(x + sign*d - px), c]
//    for d in [0, kw): the kw horizontal taps of a k x k filter become channels, so a 9x9 (or 3x3) convolution over
//    a 3-channel image turns into a kw-tap (vertical only) convolution over a 32/16-channel NHWC tensor that
//    conv_tc.cu / contract_tc.cu can consume with TMA (3 channels = 6/12 bytes per pixel cannot be a TMA row).
//    Used for: StyleTransfer first conv 9x9 3->32 (cnn.py:16) fwd + wgrad, last conv 9x9 32->3 (cnn.py:39) dgrad +
//    wgrad, VGG conv1_1 3->64 (train_cnn.py:54) fwd with the Caffe mean shift (train_cnn.py:300-301) fused in.
//  * ast_pack_weights_ex: weight re-packing with a two-level inner index and zero padding, for those layouts.
#include "common.cuh"

namespace ast {

__global__ void __launch_bounds__(256)
row_im2col_kernel(Img src, Img out, const float* __restrict__ shift, int kw, int sign, int px, int py, int reflect,
                  int round_tf32) {
  const int C = src.c;
  const long long total = (long long)out.n * out.h * out.w * out.c;
  for (long long idx = blockIdx.x * 256LL + threadIdx.x; idx < total; idx += (long long)gridDim.x * 256) {
    const int ch = (int)(idx % out.c);
    long long r = idx / out.c;
    const int x = (int)(r % out.w); r /= out.w;
    const int y = (int)(r % out.h);
    const int n = (int)(r / out.h);
    float v = 0.f;
    const int d = ch / C, c = ch - d * C;
    if (d < kw) {
      int sy = y - py, sx = x + sign * d - px;
      bool ok = true;
      if (reflect) { sy = reflect_idx(sy, src.h); sx = reflect_idx(sx, src.w); }
      else ok = sy >= 0 && sy < src.h && sx >= 0 && sx < src.w;
      if (ok) {
        v = ld_elem(src, img_off(src, n, sy, sx, c));
        if (shift) v += shift[c];
      }
    }
    if (round_tf32) { unsigned u; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v)); v = __uint_as_float(u); }
    st_elem(out, img_off(out, n, y, x, ch), v);
  }
}

// One thread per output pixel: gathers the kw*C source values (coalesced along x: consecutive threads read
// consecutive columns of each source plane) and writes the whole OUTC-channel row with 16-byte stores.
// KW, C > 0: compile-time tap / channel counts (fully unrolled, every v[] index is static); 0 = runtime values.
template <int OUTC, int KW, int CC>
__global__ void __launch_bounds__(256)
row_im2col_pix_kernel(Img src, Img out, const float* __restrict__ shift, int kw_rt, int sign, int px, int py, int reflect,
                      int round_tf32) {
  const int C = CC > 0 ? CC : src.c;
  const int kw = KW > 0 ? KW : kw_rt;
  const long long total = (long long)out.n * out.h * out.w;
  for (long long idx = blockIdx.x * 256LL + threadIdx.x; idx < total; idx += (long long)gridDim.x * 256) {
    const int x = (int)(idx % out.w);
    long long r = idx / out.w;
    const int y = (int)(r % out.h);
    const int n = (int)(r / out.h);
    float v[OUTC];
#pragma unroll
    for (int e = 0; e < OUTC; ++e) v[e] = 0.f;
    int sy = y - py;
    bool oky = true;
    if (reflect) sy = reflect_idx(sy, src.h);
    else oky = sy >= 0 && sy < src.h;
    if (oky) {
      if (KW > 0 && CC > 0) {
#pragma unroll
        for (int d = 0; d < (KW > 0 ? KW : 1); ++d) {
          int sx = x + sign * d - px;
          bool ok = true;
          if (reflect) sx = reflect_idx(sx, src.w);
          else ok = sx >= 0 && sx < src.w;
#pragma unroll
          for (int c = 0; c < (CC > 0 ? CC : 1); ++c) {
            float t = 0.f;
            if (ok) { t = ld_elem(src, img_off(src, n, sy, sx, c)); if (shift) t += shift[c]; }
            if (round_tf32) { unsigned u; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(t)); t = __uint_as_float(u); }
            if (d * CC + c < OUTC) v[d * CC + c] = t;
          }
        }
      } else {
        int ch = 0;
        for (int d = 0; d < kw; ++d) {
          int sx = x + sign * d - px;
          bool ok = true;
          if (reflect) sx = reflect_idx(sx, src.w);
          else ok = sx >= 0 && sx < src.w;
          for (int c = 0; c < C; ++c, ++ch) {
            float t = 0.f;
            if (ok) { t = ld_elem(src, img_off(src, n, sy, sx, c)); if (shift) t += shift[c]; }
            if (round_tf32) { unsigned u; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(t)); t = __uint_as_float(u); }
#pragma unroll
            for (int e = 0; e < OUTC; ++e) if (e == ch) v[e] = t;   // keeps v[] in registers
          }
        }
      }
    }
    const long long oo = img_off(out, n, y, x, 0);
    if (out.dtype == AST_F32) {
#pragma unroll
      for (int e = 0; e < OUTC; e += 4) st4((float*)out.ptr + oo + e, v + e);
    } else {
      unsigned short* op = (unsigned short*)out.ptr + oo;      // bf16 or fp16 elements: type tested once, outside the loop
      if (out.dtype == AST_F16) {
#pragma unroll
        for (int e = 0; e < OUTC; e += 8)
          *reinterpret_cast<uint4*>(op + e) = make_uint4(pack2<true>(v[e], v[e + 1]), pack2<true>(v[e + 2], v[e + 3]),
                                                         pack2<true>(v[e + 4], v[e + 5]), pack2<true>(v[e + 6], v[e + 7]));
      } else {
#pragma unroll
        for (int e = 0; e < OUTC; e += 8)
          *reinterpret_cast<uint4*>(op + e) = make_uint4(pack2<false>(v[e], v[e + 1]), pack2<false>(v[e + 2], v[e + 3]),
                                                         pack2<false>(v[e + 4], v[e + 5]), pack2<false>(v[e + 6], v[e + 7]));
      }
    }
  }
}

template <typename TO, bool ROUND_TF32>
__global__ void pack_weights_ex_kernel(const float* __restrict__ src, const int* __restrict__ tap_off, int ntaps, int a,
                                       int a_valid, int b, int b_valid, int b0, long long s_a, long long s_b1,
                                       long long s_b0, TO* __restrict__ dst) {
  const long long total = (long long)ntaps * a * b;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int ib = (int)(idx % b);
    const int ia = (int)((idx / b) % a);
    const int t = (int)(idx / ((long long)a * b));
    float v = 0.f;
    if (ia < a_valid && ib < b_valid) v = src[tap_off[t] + ia * s_a + (ib / b0) * s_b1 + (ib % b0) * s_b0];
    if (ROUND_TF32) { unsigned u; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v)); v = __uint_as_float(u); }
    DT<TO>::st(dst + idx, v);
  }
}

}  // namespace ast

namespace ast {

// Inverse companion of ast_row_im2col for a thin-OUTPUT k x k convolution (the 32->3 9x9 last layer, cnn.py:39):
// the k vertical taps run as a tcgen05 conv whose output channel r = d*C + c holds, at pixel (y, x'), the partial sum
// of output channel c for horizontal tap d; this kernel finishes the filter:
//     out[n, y, x, c] = bias[c] + sum_{d < kw} part[n, y, x + d, d*C + c]
// One block per (n, y) row: the row of partial sums is staged in shared memory with coalesced 16-byte loads.
__global__ void __launch_bounds__(256) fold_rows_kernel(Img part, Img out, const float* __restrict__ bias, int kw, int relu,
                                                        int segw) {
  extern __shared__ float row[];
  const int n = blockIdx.x / out.h, y = blockIdx.x % out.h;
  const int x0 = blockIdx.y * segw, nx = min(segw, out.w - x0);       // this block's output columns [x0, x0 + nx)
  const int pc = part.c, C = out.c;
  const float4* src = reinterpret_cast<const float4*>((const float*)part.ptr + img_off(part, n, y, x0, 0));
  const int nvec = (nx + kw - 1) * pc / 4;           // the row is contiguous (checked on the host)
  const int ps = pc + 1;                             // padded pixel stride: threads walk x, so pc (=32) would be a 32-way bank conflict
  for (int i = threadIdx.x; i < nvec; i += 256) {
    const float4 t = __ldg(src + i);
    float* dst = row + (4 * i / pc) * ps + (4 * i % pc);
    dst[0] = t.x; dst[1] = t.y; dst[2] = t.z; dst[3] = t.w;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < nx * C; idx += 256) {
    const int x = idx % nx, c = idx / nx;            // consecutive threads -> consecutive x (NCHW outputs coalesce)
    float v = bias ? bias[c] : 0.f;
    for (int d = 0; d < kw; ++d) v += row[(x + d) * ps + d * C + c];
    if (relu) v = fmaxf(v, 0.f);
    st_elem(out, img_off(out, n, y, x0 + x, c), v);
  }
}

}  // namespace ast

using namespace ast;

extern "C" int ast_fold_rows(const ast_image* part, const ast_image* out, const float* bias, int32_t kw, int32_t relu,
                             void* stream) {
  AST_CHECK_ARG(part && out, "ast_fold_rows: null argument");
  AST_CHECK_ARG(part->dtype == AST_F32 && kw >= 1 && part->n == out->n && part->h == out->h && part->w == out->w + kw - 1 &&
                part->c >= kw * out->c, "ast_fold_rows: part must be fp32 [n, h, w+kw-1, >= kw*c]");
  AST_CHECK_ARG(part->sc == 1 && part->sw == part->c && part->c % 4 == 0 && part->sh % 4 == 0 && part->sn % 4 == 0 &&
                ((uintptr_t)part->ptr & 15) == 0, "ast_fold_rows: part rows must be dense and 16-byte aligned");
  const int segw = out->w < 256 ? out->w : 256;      // output columns per block (wide rows are split into segments)
  const size_t smem = (size_t)(segw + kw - 1) * (part->c + 1) * sizeof(float);
  AST_CHECK_ARG(smem <= 200 * 1024, "ast_fold_rows: segment of %zu bytes does not fit shared memory", smem);
  if ((long long)out->n * out->h * out->w * out->c == 0) return 0;
  cudaError_t e = cudaFuncSetAttribute(fold_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("ast_fold_rows: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); return (int)e; }
  dim3 grid((unsigned)(out->n * out->h), (unsigned)((out->w + segw - 1) / segw));
  launch_k(fold_rows_kernel, grid, 256, smem, (cudaStream_t)stream, to_img(part), to_img(out), bias, kw, relu, segw);
  count_launch();
  count_work(FAM_POINTWISE, 0.0, img_bytes(part) + img_bytes(out));
  AST_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int ast_row_im2col(const ast_image* src, const ast_image* out, const float* shift, int32_t kw, int32_t sign,
                              int32_t px, int32_t py, int32_t reflect, int32_t round_tf32, void* stream) {
  AST_CHECK_ARG(src && out, "ast_row_im2col: null argument");
  AST_CHECK_ARG(out->n == src->n && kw >= 1 && out->c >= kw * src->c && (sign == 1 || sign == -1),
                "ast_row_im2col: out.c (%d) must hold kw*C (%d) channels", out->c, kw * src->c);
  AST_CHECK_ARG(!reflect || (src->h > 1 && src->w > 1), "ast_row_im2col: image too small to reflect");
  const long long total = (long long)out->n * out->h * out->w * out->c;
  if (total == 0) return 0;
  const int ev = out->dtype == AST_F32 ? 4 : 8;
  const bool vec_ok = out->sc == 1 && out->sw % ev == 0 && out->sh % ev == 0 && out->sn % ev == 0 &&
                      ((uintptr_t)out->ptr & 15) == 0 && (out->c == 16 || out->c == 32);
  if (vec_ok) {
    const long long pixels = total / out->c;
    long long pb = (pixels + 255) / 256;
    if (pb > (long long)num_sms() * 32) pb = (long long)num_sms() * 32;
#define RIK(O, K, C) launch_k(row_im2col_pix_kernel<O, K, C>, (int)pb, 256, 0, (cudaStream_t)stream, to_img(src), to_img(out), shift, kw, sign, px, py, reflect, round_tf32)
    if (out->c == 32 && kw == 9 && src->c == 3) RIK(32, 9, 3);
    else if (out->c == 16 && kw == 3 && src->c == 3) RIK(16, 3, 3);
    else if (out->c == 32 && kw == 3 && src->c == 3) RIK(32, 3, 3);
    else if (out->c == 32) RIK(32, 0, 0);
    else RIK(16, 0, 0);
#undef RIK
    count_launch();
    count_work(FAM_POINTWISE, 0.0, img_bytes(src) + img_bytes(out));
    AST_CUDA_LAUNCH_CHECK();
    return 0;
  }
  long long blocks = (total + 255) / 256;
  if (blocks > (long long)num_sms() * 16) blocks = (long long)num_sms() * 16;
  launch_k(row_im2col_kernel, (int)blocks, 256, 0, (cudaStream_t)stream, to_img(src), to_img(out), shift, kw, sign, px, py,
                                                                   reflect, round_tf32);
  count_launch();
  count_work(FAM_POINTWISE, 0.0, img_bytes(src) + img_bytes(out));
  AST_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int ast_pack_weights_ex(const float* src, const int32_t* tap_off, int32_t ntaps, int32_t a, int32_t a_valid,
                                   int32_t b, int32_t b_valid, int32_t b0, int64_t s_a, int64_t s_b1, int64_t s_b0,
                                   void* dst, int32_t dst_dtype, void* stream) {
  AST_CHECK_ARG(src && tap_off && dst, "ast_pack_weights_ex: null argument");
  AST_CHECK_ARG(ntaps >= 1 && ntaps <= AST_MAX_TAPS && a > 0 && b > 0 && b0 > 0 && a_valid <= a && b_valid <= b,
                "ast_pack_weights_ex: bad sizes");
  const long long total = (long long)ntaps * a * b;
  long long blocks = (total + 255) / 256;
  if (blocks > (long long)num_sms() * 8) blocks = (long long)num_sms() * 8;
  cudaStream_t s = (cudaStream_t)stream;
#define PK(T, R) launch_k(pack_weights_ex_kernel<T, R>, (int)blocks, 256, 0, s, src, tap_off, ntaps, a, a_valid, b, b_valid, b0, s_a, s_b1, s_b0, (T*)dst)
  if (dst_dtype == AST_F32) PK(float, false);
  else if (dst_dtype == AST_TF32) PK(float, true);
  else if (dst_dtype == AST_BF16) PK(__nv_bfloat16, false);
  else AST_CHECK_ARG(false, "ast_pack_weights_ex: bad dtype %d", dst_dtype);
#undef PK
  count_launch();
  count_work(FAM_OPTIM, 0.0, 6.0 * total);
  AST_CUDA_LAUNCH_CHECK();
  return 0;
}
