// SIMT (FFMA, fp32-accumulate) gather convolution, its weight gradient and the weight packer.
// This is the strict-fp32 precision mode of the path (north_star: 1e-5 on Grams/losses) and the
// on-device cross-check for the tcgen05 kernels in conv_tc.cu; it is selected explicitly, never as a
// fallback.  Replaces aten::mkldnn_convolution / cudnn conv, conv_transpose2d, convolution_backward
// at the call sites listed in include/ast.h.
#include "common.cuh"

namespace ast {

struct GeomDev {
  int mi, mj, si, so, oy0, ox0, ntaps, flags;
  long long w_img_stride;
  short dy[AST_MAX_TAPS];
  short dx[AST_MAX_TAPS];
};

static GeomDev to_dev(const ast_gather_geom* g) {
  GeomDev d;
  d.mi = g->mi; d.mj = g->mj; d.si = g->si; d.so = g->so; d.oy0 = g->oy0; d.ox0 = g->ox0;
  d.ntaps = g->ntaps; d.flags = g->flags; d.w_img_stride = g->w_img_stride;
  for (int t = 0; t < AST_MAX_TAPS; ++t) { d.dy[t] = g->dy[t]; d.dx[t] = g->dx[t]; }
  return d;
}

constexpr int BM = 64, BN = 64, BK = 16;

template <typename TI>
__global__ void __launch_bounds__(256)
conv_gather_simt_kernel(Img in, const TI* __restrict__ wts, const float* __restrict__ bias,
                        const float* __restrict__ in_shift, Img add, Img mask, Img out, GeomDev g) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];

  const int tid = threadIdx.x;
  const int n = blockIdx.z;
  const int m0 = blockIdx.x * BM;
  const int co0 = blockIdx.y * BN;
  const int cin = in.c, cout = out.c;
  const int mtot = g.mi * g.mj;
  const TI* w = wts + (long long)n * g.w_img_stride;
  const TI* inp = (const TI*)in.ptr;

  // loader role: one pixel (lp) and one channel quad (lq) of the A tile; one cout (lp) and quad of B
  const int lp = tid >> 2, lq = tid & 3;
  const int lm = m0 + lp;
  const bool lvalid = lm < mtot;
  const int li = lvalid ? lm / g.mj : 0, lj = lvalid ? lm % g.mj : 0;
  const int iy0 = g.si * li, ix0 = g.si * lj;
  const bool reflect = g.flags & AST_CONV_REFLECT;
  const bool vecA = (in.sc == 1) && ((cin & 3) == 0) && ((in.sw & 3) == 0) && ((in.sh & 3) == 0) && ((in.sn & 3) == 0);
  const bool vecB = (cin & 3) == 0;

  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;

  for (int t = 0; t < g.ntaps; ++t) {
    int y = iy0 + g.dy[t], x = ix0 + g.dx[t];
    bool ok = lvalid;
    if (reflect) { y = reflect_idx(y, in.h); x = reflect_idx(x, in.w); }
    else ok = ok && y >= 0 && y < in.h && x >= 0 && x < in.w;
    const long long abase = ok ? img_off(in, n, y, x, 0) : 0;
    const TI* wt = w + (long long)t * cout * cin;
    for (int k0 = 0; k0 < cin; k0 += BK) {
      // ---- A tile: As[k][pixel]
      {
        const int c = k0 + lq * 4;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (ok) {
          if (vecA && c + 3 < cin) {
            if (sizeof(TI) == 4) {
              float4 f = *reinterpret_cast<const float4*>((const float*)inp + abase + c);
              v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
            } else {
              uint2 u = *reinterpret_cast<const uint2*>((const __nv_bfloat16*)inp + abase + c);
              const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
              float2 f0 = __bfloat1622float2(h[0]), f1 = __bfloat1622float2(h[1]);
              v[0] = f0.x; v[1] = f0.y; v[2] = f1.x; v[3] = f1.y;
            }
            if (in_shift) {
#pragma unroll
              for (int e = 0; e < 4; ++e) v[e] += in_shift[c + e];
            }
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (c + e < cin) {
                v[e] = DT<TI>::ld(inp + abase + (long long)(c + e) * in.sc);
                if (in_shift) v[e] += in_shift[c + e];
              }
          }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) As[lq * 4 + e][lp] = v[e];
      }
      // ---- B tile: Bs[k][cout]
      {
        const int co = co0 + lp;
        const int c = k0 + lq * 4;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (co < cout) {
          const TI* p = wt + (long long)co * cin + c;
          if (vecB && c + 3 < cin) {
            if (sizeof(TI) == 4) {
              float4 f = *reinterpret_cast<const float4*>(p);
              v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
            } else {
              uint2 u = *reinterpret_cast<const uint2*>(p);
              const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
              float2 f0 = __bfloat1622float2(h[0]), f1 = __bfloat1622float2(h[1]);
              v[0] = f0.x; v[1] = f0.y; v[2] = f1.x; v[3] = f1.y;
            }
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (c + e < cin) v[e] = DT<TI>::ld(p + e);
          }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) Bs[lq * 4 + e][lp] = v[e];
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
        const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
        const float a[4] = {a4.x, a4.y, a4.z, a4.w};
        const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[p][q] = fmaf(a[p], b[q], acc[p][q]);
      }
      __syncthreads();
    }
  }

  // ---- epilogue: bias, add, relu, mask, strided store
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int m = m0 + ty * 4 + p;
    if (m >= mtot) continue;
    const int i = m / g.mj, j = m % g.mj;
    const int oy = g.oy0 + g.so * i, ox = g.ox0 + g.so * j;
    if (oy >= out.h || ox >= out.w) continue;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int co = co0 + tx * 4 + q;
      if (co >= cout) continue;
      float v = acc[p][q];
      if (bias) v += bias[co];
      if (add.ptr) v += ld_elem(add, img_off(add, n, oy, ox, co));
      if (g.flags & AST_CONV_RELU) v = fmaxf(v, 0.f);
      if (mask.ptr) v = ld_elem(mask, img_off(mask, n, oy, ox, co)) > 0.f ? v : 0.f;
      if (g.flags & AST_CONV_ROUND_TF32) { unsigned r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v)); v = __uint_as_float(r); }
      st_elem(out, img_off(out, n, oy, ox, co), v);
    }
  }
}

// ---- weight gradient: dW_t[co][ci] = sum_pixels gout[p][co] * x[p_t][ci]
template <typename TX, typename TG>
__global__ void __launch_bounds__(256)
wgrad_gather_simt_kernel(Img x, Img gout, float* __restrict__ dw, const int* __restrict__ tap_off,
                         long long s_co, long long s_ci, GeomDev g, int nci_tiles, long long chunk,
                         long long dw_img_stride, int ksplit, float scale) {
  __shared__ __align__(16) float Ys[BK][BN + 4];
  __shared__ __align__(16) float Xs[BK][BM + 4];
  const int tid = threadIdx.x;
  const int t = blockIdx.y;
  const int co0 = (blockIdx.z / nci_tiles) * BN;
  const int ci0 = (blockIdx.z % nci_tiles) * BM;
  const int cout = gout.c, cin = x.c;
  const long long per_img = (long long)g.mi * g.mj;
  long long mbeg, mend;
  if (dw_img_stride) {  // per-image outputs (Gram): blockIdx.x = image * ksplit + chunk index
    const int img = blockIdx.x / ksplit, kc = blockIdx.x % ksplit;
    mbeg = img * per_img + (long long)kc * chunk;
    mend = min((img + 1) * per_img, mbeg + chunk);
    dw += img * dw_img_stride;
  } else {
    mbeg = (long long)blockIdx.x * chunk;
    mend = min(per_img * x.n, mbeg + chunk);
  }
  const bool reflect = g.flags & AST_CONV_REFLECT;
  const int dyt = g.dy[t], dxt = g.dx[t];

  const int lp = tid >> 4, lq = tid & 15;  // pixel within BK, channel quad
  const int ty = tid >> 4, tx = tid & 15;  // co quad, ci quad
  const bool vecY = gout.sc == 1 && (cout & 3) == 0 && (gout.sw & 3) == 0 && (gout.sh & 3) == 0 && (gout.sn & 3) == 0;
  const bool vecX = x.sc == 1 && (cin & 3) == 0 && (x.sw & 3) == 0 && (x.sh & 3) == 0 && (x.sn & 3) == 0;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;

  for (long long mk = mbeg; mk < mend; mk += BK) {
    const long long m = mk + lp;
    float yv[4] = {0.f, 0.f, 0.f, 0.f}, xv[4] = {0.f, 0.f, 0.f, 0.f};
    if (m < mend) {
      const int n = (int)(m / per_img);
      const int r = (int)(m % per_img);
      const int i = r / g.mj, j = r % g.mj;
      const int oy = g.oy0 + g.so * i, ox = g.ox0 + g.so * j;
      int iy = g.si * i + dyt, ix = g.si * j + dxt;
      bool okx = true;
      if (reflect) { iy = reflect_idx(iy, x.h); ix = reflect_idx(ix, x.w); }
      else okx = iy >= 0 && iy < x.h && ix >= 0 && ix < x.w;
      const bool oky = oy < gout.h && ox < gout.w;
      if (oky && okx) {
        const int cy = co0 + lq * 4, cx = ci0 + lq * 4;
        const long long yb = img_off(gout, n, oy, ox, 0), xb = img_off(x, n, iy, ix, 0);
        if (vecY && cy + 3 < cout) {
          if (sizeof(TG) == 4) { float4 f = *reinterpret_cast<const float4*>((const float*)gout.ptr + yb + cy); yv[0] = f.x; yv[1] = f.y; yv[2] = f.z; yv[3] = f.w; }
          else {
            uint2 u = *reinterpret_cast<const uint2*>((const __nv_bfloat16*)gout.ptr + yb + cy);
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
            float2 f0 = __bfloat1622float2(h[0]), f1 = __bfloat1622float2(h[1]);
            yv[0] = f0.x; yv[1] = f0.y; yv[2] = f1.x; yv[3] = f1.y;
          }
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) if (cy + e < cout) yv[e] = DT<TG>::ld((const TG*)gout.ptr + yb + (long long)(cy + e) * gout.sc);
        }
        if (vecX && cx + 3 < cin) {
          if (sizeof(TX) == 4) { float4 f = *reinterpret_cast<const float4*>((const float*)x.ptr + xb + cx); xv[0] = f.x; xv[1] = f.y; xv[2] = f.z; xv[3] = f.w; }
          else {
            uint2 u = *reinterpret_cast<const uint2*>((const __nv_bfloat16*)x.ptr + xb + cx);
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
            float2 f0 = __bfloat1622float2(h[0]), f1 = __bfloat1622float2(h[1]);
            xv[0] = f0.x; xv[1] = f0.y; xv[2] = f1.x; xv[3] = f1.y;
          }
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) if (cx + e < cin) xv[e] = DT<TX>::ld((const TX*)x.ptr + xb + (long long)(cx + e) * x.sc);
        }
      }
    }
    *reinterpret_cast<float4*>(&Ys[lp][lq * 4]) = make_float4(yv[0], yv[1], yv[2], yv[3]);
    *reinterpret_cast<float4*>(&Xs[lp][lq * 4]) = make_float4(xv[0], xv[1], xv[2], xv[3]);
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&Ys[kk][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Xs[kk][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[p][q] = fmaf(a[p], b[q], acc[p][q]);
    }
    __syncthreads();
  }
  const long long toff = tap_off ? tap_off[t] : 0;
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int co = co0 + ty * 4 + p;
    if (co >= cout) continue;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int ci = ci0 + tx * 4 + q;
      if (ci >= cin) continue;
      atomicAdd(dw + toff + co * s_co + ci * s_ci, acc[p][q] * scale);
    }
  }
}

template <typename TO, bool ROUND_TF32 = false>
__global__ void pack_weights_kernel(const float* __restrict__ src, const int* __restrict__ tap_off, int ntaps,
                                    int a, int b, long long s_a, long long s_b, TO* __restrict__ dst) {
  const long long total = (long long)ntaps * a * b;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int ib = (int)(idx % b);
    const int ia = (int)((idx / b) % a);
    const int t = (int)(idx / ((long long)a * b));
    float v = src[tap_off[t] + ia * s_a + ib * s_b];
    if (ROUND_TF32) { unsigned r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v)); v = __uint_as_float(r); }
    DT<TO>::st(dst + idx, v);
  }
}

int conv_gather_tc(const ast_image* in, const void* weights, const float* bias, const float* in_shift,
                   const ast_image* add, const ast_image* mask, const ast_image* out,
                   const ast_gather_geom* geom, cudaStream_t stream);  // conv_tc.cu
struct GramFin;
int contract_tc(const ast_image* rows, int r_s, int r_oy, int r_ox, const ast_image* cols, int c_s, const short* dy,
                const short* dx, int ntaps, int mi, int mj, float* out, const int* tap_off, long long s_m,
                long long s_n, long long out_img_stride, float scale, int upper_only, cudaStream_t stream,
                const GramFin* finp);  // contract_tc.cu

}  // namespace ast

using namespace ast;

extern "C" int ast_conv_gather(const ast_image* in, const void* weights, const float* bias, const float* in_shift,
                               const ast_image* add, const ast_image* mask, const ast_image* out,
                               const ast_gather_geom* geom, void* stream) {
  AST_CHECK_ARG(in && weights && out && geom, "ast_conv_gather: null argument");
  AST_CHECK_ARG(geom->ntaps >= 1 && geom->ntaps <= AST_MAX_TAPS, "ast_conv_gather: ntaps %d out of range", geom->ntaps);
  AST_CHECK_ARG(in->n == out->n, "ast_conv_gather: batch mismatch %d vs %d", in->n, out->n);
  AST_CHECK_ARG(geom->mi > 0 && geom->mj > 0 && geom->si >= 1 && geom->so >= 1, "ast_conv_gather: bad geometry");
  AST_CHECK_ARG(!add || same_shape(add, out), "ast_conv_gather: add image shape mismatch");
  AST_CHECK_ARG(!mask || same_shape(mask, out), "ast_conv_gather: mask image shape mismatch");
  AST_CHECK_ARG(in->dtype == AST_F32 || in->dtype == AST_BF16 || (in->dtype == AST_F16 && (geom->flags & AST_CONV_TENSOR)),
                "ast_conv_gather: bad input dtype (fp16 operands are a tensor-core feature)");
  AST_CHECK_ARG(!geom->stats || (geom->flags & AST_CONV_TENSOR), "ast_conv_gather: fused statistics are a tensor-core epilogue feature");
  AST_CHECK_ARG(!geom->pooled || (geom->flags & AST_CONV_TENSOR), "ast_conv_gather: the pooled output is a tensor-core epilogue feature");
  if (geom->flags & AST_CONV_TENSOR)
    return conv_gather_tc(in, weights, bias, in_shift, add, mask, out, geom, (cudaStream_t)stream);
  if (in->n == 0 || out->c == 0) return 0;
  GeomDev g = to_dev(geom);
  const int mtot = geom->mi * geom->mj;
  dim3 grid((mtot + BM - 1) / BM, (out->c + BN - 1) / BN, in->n);
  Img addi = add ? to_img(add) : null_img(), maski = mask ? to_img(mask) : null_img();
  if (in->dtype == AST_F32)
    launch_k(conv_gather_simt_kernel<float>, grid, 256, 0, (cudaStream_t)stream, 
        to_img(in), (const float*)weights, bias, in_shift, addi, maski, to_img(out), g);
  else
    launch_k(conv_gather_simt_kernel<__nv_bfloat16>, grid, 256, 0, (cudaStream_t)stream, 
        to_img(in), (const __nv_bfloat16*)weights, bias, in_shift, addi, maski, to_img(out), g);
  count_launch();
  count_work(FAM_CONV_SIMT, conv_flops(in, out, geom), conv_bytes(in, out, geom, add, mask));
  AST_CUDA_LAUNCH_CHECK();
  return 0;
}

namespace ast {
int launch_wgrad_simt(const ast_image* x, const ast_image* gout, float* dw, const int32_t* tap_off,
                      int64_t s_co, int64_t s_ci, const ast_gather_geom* geom, int64_t dw_img_stride, float scale,
                      cudaStream_t s) {
  GeomDev g = to_dev(geom);
  const int nco = (gout->c + BN - 1) / BN, nci = (x->c + BM - 1) / BM;
  const long long per_img = (long long)geom->mi * geom->mj;
  const long long mtot = dw_img_stride ? per_img : per_img * x->n;   // reduction length per output
  const long long fixed = (long long)geom->ntaps * nco * nci * (dw_img_stride ? x->n : 1);
  long long ksplit = (4LL * num_sms() + fixed - 1) / fixed;
  if (ksplit < 1) ksplit = 1;
  long long chunk = (mtot + ksplit - 1) / ksplit;
  chunk = ((chunk + BK - 1) / BK) * BK;
  if (chunk < 256) chunk = 256;
  ksplit = (mtot + chunk - 1) / chunk;
  dim3 grid((unsigned)(ksplit * (dw_img_stride ? x->n : 1)), geom->ntaps, nco * nci);
#define LAUNCH(TX, TG) launch_k(wgrad_gather_simt_kernel<TX, TG>, grid, 256, 0, s, to_img(x), to_img(gout), dw, tap_off, s_co, s_ci, g, nci, chunk, dw_img_stride, (int)ksplit, scale)
  if (x->dtype == AST_F32 && gout->dtype == AST_F32) LAUNCH(float, float);
  else if (x->dtype == AST_BF16 && gout->dtype == AST_BF16) LAUNCH(__nv_bfloat16, __nv_bfloat16);
  else if (x->dtype == AST_F32 && gout->dtype == AST_BF16) LAUNCH(float, __nv_bfloat16);
  else LAUNCH(__nv_bfloat16, float);
#undef LAUNCH
  count_launch();
  if (dw_img_stride)
    count_work(FAM_GRAM_SIMT, (double)x->n * x->c * (x->c + 1.0) * per_img, img_bytes(x) + 4.0 * x->n * x->c * x->c);
  else
    count_work(FAM_WGRAD_SIMT, 2.0 * x->n * per_img * geom->ntaps * x->c * gout->c, img_bytes(x) + img_bytes(gout));
  AST_CUDA_LAUNCH_CHECK();
  return 0;
}
}  // namespace ast

extern "C" int ast_wgrad_gather(const ast_image* x, const ast_image* gout, float* dw, const int32_t* tap_off,
                                int64_t s_co, int64_t s_ci, const ast_gather_geom* geom, void* stream) {
  AST_CHECK_ARG(x && gout && dw && tap_off && geom, "ast_wgrad_gather: null argument");
  AST_CHECK_ARG(geom->ntaps >= 1 && geom->ntaps <= AST_MAX_TAPS, "ast_wgrad_gather: ntaps %d out of range", geom->ntaps);
  AST_CHECK_ARG(x->n == gout->n, "ast_wgrad_gather: batch mismatch");
  if (x->n == 0) return 0;
  if (geom->flags & AST_CONV_TENSOR)
    return contract_tc(gout, geom->so, geom->oy0, geom->ox0, x, geom->si, geom->dy, geom->dx, geom->ntaps, geom->mi,
                       geom->mj, dw, tap_off, s_co, s_ci, 0, 1.f, 0, (cudaStream_t)stream, nullptr);
  return launch_wgrad_simt(x, gout, dw, tap_off, s_co, s_ci, geom, 0, 1.f, (cudaStream_t)stream);
}

extern "C" int ast_pack_weights(const float* src, const int32_t* tap_off, int32_t ntaps, int32_t a, int32_t b,
                                int64_t s_a, int64_t s_b, void* dst, int32_t dst_dtype, void* stream) {
  AST_CHECK_ARG(src && tap_off && dst, "ast_pack_weights: null argument");
  AST_CHECK_ARG(ntaps >= 1 && ntaps <= AST_MAX_TAPS && a > 0 && b > 0, "ast_pack_weights: bad sizes");
  const long long total = (long long)ntaps * a * b;
  const int blocks = (int)min((total + 255) / 256, (long long)num_sms() * 8);
  if (dst_dtype == AST_F32)
    launch_k(pack_weights_kernel<float>, blocks, 256, 0, (cudaStream_t)stream, src, tap_off, ntaps, a, b, s_a, s_b, (float*)dst);
  else if (dst_dtype == AST_TF32)
    launch_k(pack_weights_kernel<float, true>, blocks, 256, 0, (cudaStream_t)stream, src, tap_off, ntaps, a, b, s_a, s_b, (float*)dst);
  else if (dst_dtype == AST_BF16)
    launch_k(pack_weights_kernel<__nv_bfloat16>, blocks, 256, 0, (cudaStream_t)stream, src, tap_off, ntaps, a, b, s_a, s_b, (__nv_bfloat16*)dst);
  else AST_CHECK_ARG(false, "ast_pack_weights: bad dtype %d", dst_dtype);
  count_launch();
  AST_CUDA_LAUNCH_CHECK();
  return 0;
}
