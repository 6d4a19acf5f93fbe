// Fused InstanceNorm kernels (HBM-bound, NHWC, 16-byte vector accesses, warp/block reductions).
// Replaces aten::instance_norm -> native_batch_norm, aten::relu, the residual aten::add and
// aten::reflection_pad2d at cnn.py:58,68,78,96-98,114,123 and their autograd backward.
#include "common.cuh"

namespace ast {

constexpr int NT = 256;

// ---------------------------------------------------------------- forward statistics
// Per-thread shifted sums -> (count, mean, M2); Chan merge inside the block and across blocks.
template <typename T>
__global__ void __launch_bounds__(NT)
in_stats_partial_kernel(Img x, float* __restrict__ ws, int nblk, int chunk) {
  constexpr int VEC = Vec16<T>::N;
  extern __shared__ float sm[];
  const int C = x.c, lanes = C / VEC, slots = NT / lanes;
  const int tid = threadIdx.x, lane = tid % lanes, slot = tid / lanes;
  const int n = blockIdx.y, blk = blockIdx.x;
  const int hw = x.h * x.w;
  const int pbeg = blk * chunk, pend = min(hw, pbeg + chunk);
  const T* base = (const T*)x.ptr + (long long)n * x.sn + lane * VEC;

  float k[VEC], s1[VEC], s2[VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) { k[e] = 0.f; s1[e] = 0.f; s2[e] = 0.f; }
  int cnt = 0;
  if (slot < slots) {
    for (int p = pbeg + slot; p < pend; p += slots) {
      const int y = p / x.w, xx = p - y * x.w;
      float v[VEC];
      Vec16<T>::load(base + (long long)y * x.sh + (long long)xx * x.sw, v);
      if (cnt == 0) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) k[e] = v[e];
      }
#pragma unroll
      for (int e = 0; e < VEC; ++e) { const float d = v[e] - k[e]; s1[e] += d; s2[e] = fmaf(d, d, s2[e]); }
      ++cnt;
    }
  }
  // smem: cntS[slots] | meanS[slots][C] | m2S[slots][C]
  float* cntS = sm;
  float* meanS = sm + NT;
  float* m2S = meanS + slots * C;
  if (slot < slots) {
    if (lane == 0) cntS[slot] = (float)cnt;
    const float inv = cnt > 0 ? 1.f / cnt : 0.f;
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      meanS[slot * C + lane * VEC + e] = k[e] + s1[e] * inv;
      m2S[slot * C + lane * VEC + e] = fmaxf(s2[e] - s1[e] * s1[e] * inv, 0.f);
    }
  }
  __syncthreads();
  for (int c = tid; c < C; c += NT) {
    float na = 0.f, ma = 0.f, qa = 0.f;
    for (int s = 0; s < slots; ++s) {
      const float nb = cntS[s];
      if (nb == 0.f) continue;
      const float mb = meanS[s * C + c], qb = m2S[s * C + c];
      const float nab = na + nb, d = mb - ma;
      ma += d * (nb / nab);
      qa += qb + d * d * (na * nb / nab);
      na = nab;
    }
    float* o = ws + ((long long)(n * nblk + blk) * C + c) * 3;
    o[0] = na; o[1] = ma; o[2] = qa;
  }
}

__global__ void in_stats_final_kernel(const float* __restrict__ ws, int nblk, int C, int hw, float eps,
                                      float* __restrict__ mean, float* __restrict__ rstd) {
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float na = 0.f, ma = 0.f, qa = 0.f;
    for (int b = 0; b < nblk; ++b) {
      const float* o = ws + ((long long)(n * nblk + b) * C + c) * 3;
      const float nb = o[0];
      if (nb == 0.f) continue;
      const float d = o[1] - ma, nab = na + nb;
      ma += d * (nb / nab);
      qa += o[2] + d * d * (na * nb / nab);
      na = nab;
    }
    mean[n * C + c] = ma;
    rstd[n * C + c] = rsqrtf(qa / (float)hw + eps);
  }
}

// ---------------------------------------------------------------- forward apply (+ReLU, +residual, reflect-pad write)
template <typename TX, typename TO>
__global__ void __launch_bounds__(NT)
in_apply_kernel(Img x, const float* __restrict__ mean, const float* __restrict__ rstd,
                const float* __restrict__ gamma, const float* __restrict__ beta, Img res, Img out, int pad, int relu) {
  constexpr int VEC = 4;  // 4 channels per thread regardless of dtype (8B bf16 / 16B fp32 accesses)
  const int C = x.c, lanes = C / VEC;
  const long long total = (long long)out.n * out.h * out.w * lanes;
  for (long long idx = blockIdx.x * (long long)NT + threadIdx.x; idx < total; idx += (long long)gridDim.x * NT) {
    const int lane = (int)(idx % lanes);
    long long r = idx / lanes;
    const int ox = (int)(r % out.w); r /= out.w;
    const int oy = (int)(r % out.h);
    const int n = (int)(r / out.h);
    const int i = reflect_idx(oy - pad, x.h), j = reflect_idx(ox - pad, x.w);
    const int c = lane * VEC;
    const long long xo = img_off(x, n, i, j, c);
    float v[VEC];
    ld4((const TX*)x.ptr + xo, v);
    const int sc = n * C + c;
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const float a = gamma[c + e] * rstd[sc + e];
      v[e] = fmaf(a, v[e] - mean[sc + e], beta[c + e]);
    }
    if (res.ptr) {
      float rv[VEC];
      ld4_img(res, img_off(res, n, i, j, c), rv);
#pragma unroll
      for (int e = 0; e < VEC; ++e) v[e] += rv[e];
    }
    if (relu) {
#pragma unroll
      for (int e = 0; e < VEC; ++e) v[e] = fmaxf(v[e], 0.f);
    }
    st4((TO*)out.ptr + img_off(out, n, oy, ox, c), v);
  }
}

// ---------------------------------------------------------------- backward
// g' at interior pixel (i,j): folded padded gradient + extra gradient, masked by ReLU(y) if requested.
struct FoldIdx { int r[3]; int nr; };
__device__ __forceinline__ FoldIdx fold_rows(int i, int h, int pad) {
  FoldIdx f; f.nr = 0;
  f.r[f.nr++] = i + pad;
  if (pad > 0) {
    if (i >= 1 && i <= pad) f.r[f.nr++] = pad - i;
    if (i <= h - 2 && i >= h - 1 - pad) f.r[f.nr++] = pad + 2 * (h - 1) - i;
  }
  return f;
}

template <typename TX, int VEC>
__device__ __forceinline__ void in_bwd_gprime(const Img& x, const Img& gpad, int pad, const Img& gextra, int relu,
                                              const float* yv, int n, int i, int j, int c, float* g) {
#pragma unroll
  for (int e = 0; e < VEC; ++e) g[e] = 0.f;
  if (gpad.ptr) {
    const FoldIdx fr = fold_rows(i, x.h, pad), fc = fold_rows(j, x.w, pad);
    for (int a_ = 0; a_ < fr.nr; ++a_)
      for (int b_ = 0; b_ < fc.nr; ++b_) {
        float t[VEC];
        ld4_img(gpad, img_off(gpad, n, fr.r[a_], fc.r[b_], c), t);
#pragma unroll
        for (int e = 0; e < VEC; ++e) g[e] += t[e];
      }
  }
  if (gextra.ptr) {
    float t[VEC];
    ld4_img(gextra, img_off(gextra, n, i, j, c), t);
#pragma unroll
    for (int e = 0; e < VEC; ++e) g[e] += t[e];
  }
  if (relu) {
#pragma unroll
    for (int e = 0; e < VEC; ++e) g[e] = yv[e] > 0.f ? g[e] : 0.f;
  }
}

template <typename TX>
__global__ void __launch_bounds__(NT)
in_bwd_stats_kernel(Img x, const float* __restrict__ mean, const float* __restrict__ rstd,
                    const float* __restrict__ gamma, const float* __restrict__ beta, Img gpad, int pad, Img gextra,
                    int relu, float* __restrict__ s1o, float* __restrict__ s2o, int chunk) {
  constexpr int VEC = 4;
  extern __shared__ float sm[];
  const int C = x.c, lanes = C / VEC, slots = NT / lanes;
  const int tid = threadIdx.x, lane = tid % lanes, slot = tid / lanes;
  const int n = blockIdx.y;
  const int hw = x.h * x.w;
  const int pbeg = blockIdx.x * chunk, pend = min(hw, pbeg + chunk);
  const int c = lane * VEC;
  float a[VEC], b[VEC], mu[VEC], rs[VEC], t1[VEC], t2[VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) {
    mu[e] = mean[n * C + c + e]; rs[e] = rstd[n * C + c + e];
    a[e] = gamma[c + e] * rs[e]; b[e] = beta[c + e];
    t1[e] = 0.f; t2[e] = 0.f;
  }
  if (slot < slots) {
    for (int p = pbeg + slot; p < pend; p += slots) {
      const int i = p / x.w, j = p - i * x.w;
      float xv[VEC], g[VEC];
      ld4((const TX*)x.ptr + img_off(x, n, i, j, c), xv);
      float yv[VEC];
#pragma unroll
      for (int e = 0; e < VEC; ++e) yv[e] = fmaf(a[e], xv[e] - mu[e], b[e]);   // same expression as in_apply_kernel
      in_bwd_gprime<TX, VEC>(x, gpad, pad, gextra, relu, yv, n, i, j, c, g);
#pragma unroll
      for (int e = 0; e < VEC; ++e) { t1[e] += g[e]; t2[e] = fmaf(g[e], (xv[e] - mu[e]) * rs[e], t2[e]); }
    }
  }
  float* r1 = sm;               // [slots][C]
  float* r2 = sm + slots * C;
  if (slot < slots) {
#pragma unroll
    for (int e = 0; e < VEC; ++e) { r1[slot * C + c + e] = t1[e]; r2[slot * C + c + e] = t2[e]; }
  }
  __syncthreads();
  for (int cc = tid; cc < C; cc += NT) {
    float u1 = 0.f, u2 = 0.f;
    for (int s = 0; s < slots; ++s) { u1 += r1[s * C + cc]; u2 += r2[s * C + cc]; }
    atomicAdd(s1o + n * C + cc, u1);
    atomicAdd(s2o + n * C + cc, u2);
  }
}

template <typename TX, typename TO>
__global__ void __launch_bounds__(NT)
in_bwd_apply_kernel(Img x, const float* __restrict__ mean, const float* __restrict__ rstd,
                    const float* __restrict__ gamma, const float* __restrict__ beta, Img gpad, int pad, Img gextra,
                    int relu, const float* __restrict__ s1, const float* __restrict__ s2, Img dx, Img gtotal) {
  constexpr int VEC = 4;
  const int C = x.c, lanes = C / VEC;
  const float inv_hw = 1.f / (float)(x.h * x.w);
  const long long total = (long long)x.n * x.h * x.w * lanes;
  for (long long idx = blockIdx.x * (long long)NT + threadIdx.x; idx < total; idx += (long long)gridDim.x * NT) {
    const int lane = (int)(idx % lanes);
    long long r = idx / lanes;
    const int j = (int)(r % x.w); r /= x.w;
    const int i = (int)(r % x.h);
    const int n = (int)(r / x.h);
    const int c = lane * VEC;
    float a[VEC], b[VEC], mu[VEC], rs[VEC], xv[VEC], g[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      mu[e] = mean[n * C + c + e]; rs[e] = rstd[n * C + c + e];
      a[e] = gamma[c + e] * rs[e]; b[e] = beta[c + e];
    }
    ld4((const TX*)x.ptr + img_off(x, n, i, j, c), xv);
    float yv[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) yv[e] = fmaf(a[e], xv[e] - mu[e], b[e]);
    in_bwd_gprime<TX, VEC>(x, gpad, pad, gextra, relu, yv, n, i, j, c, g);
    float d[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const float xh = (xv[e] - mu[e]) * rs[e];
      d[e] = a[e] * (g[e] - s1[n * C + c + e] * inv_hw - xh * s2[n * C + c + e] * inv_hw);
    }
    st4((TO*)dx.ptr + img_off(dx, n, i, j, c), d);
    if (gtotal.ptr) st4_img(gtotal, img_off(gtotal, n, i, j, c), g);
  }
}

__global__ void in_finalize_kernel(const double* __restrict__ sums, int total, double inv_hw, float eps,
                                   float* __restrict__ mean, float* __restrict__ rstd) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const double m = sums[2 * i] * inv_hw;
  const double var = fmax(sums[2 * i + 1] * inv_hw - m * m, 0.0);
  mean[i] = (float)m;
  rstd[i] = (float)(1.0 / sqrt(var + (double)eps));
}

static int stats_blocks(int n, int hw, int slots) {
  int nblk = (4 * num_sms() + n - 1) / n;
  const int maxb = (hw + slots * 4 - 1) / (slots * 4);
  if (nblk > maxb) nblk = maxb;
  if (nblk < 1) nblk = 1;
  return nblk;
}

static bool nhwc_ok(const ast_image* x, int vec) {
  return x->sc == 1 && x->c % vec == 0 && x->sw % vec == 0 && x->sh % vec == 0 && x->sn % vec == 0;
}

}  // namespace ast

namespace ast {   // norm_staged.cu: return 1 = launched, 0 = not applicable (use the generic kernel), other = error
int instnorm_apply_staged(const ast_image* x, const float* mean, const float* rstd, const float* gamma, const float* beta,
                          const ast_image* residual, const ast_image* out, int pad, int relu, cudaStream_t s);
int instnorm_bwd_stats_staged(const ast_image* x, const float* mean, const float* rstd, const float* gamma,
                              const float* beta, const ast_image* gpad, int pad, const ast_image* gextra, int relu,
                              float* s1, float* s2, cudaStream_t s);
int instnorm_bwd_apply_staged(const ast_image* x, const float* mean, const float* rstd, const float* gamma,
                              const float* beta, const ast_image* gpad, int pad, const ast_image* gextra, int relu,
                              const float* s1, const float* s2, const ast_image* dx, const ast_image* gtotal,
                              cudaStream_t s);
int instnorm_bwd_fused_staged(const ast_image* x, const float* mean, const float* rstd, const float* gamma,
                              const float* beta, const ast_image* gpad, int pad, const ast_image* gextra, int relu,
                              float* s1, float* s2, int* arrive, const ast_image* dx, const ast_image* gtotal,
                              cudaStream_t s);
}  // namespace ast

using namespace ast;

extern "C" int64_t ast_instnorm_workspace_bytes(int32_t n, int32_t c) {
  return (int64_t)(4 * 148 * 2 + n) * c * 3 * sizeof(float);
}

extern "C" int ast_instnorm_stats(const ast_image* x, float* mean, float* rstd, float eps, void* workspace, void* stream) {
  AST_CHECK_ARG(x && mean && rstd && workspace, "ast_instnorm_stats: null argument");
  const int vec = x->dtype == AST_F32 ? 4 : 8;
  AST_CHECK_ARG(nhwc_ok(x, vec), "ast_instnorm_stats: needs NHWC with C %% %d == 0 (c=%d sc=%lld)", vec, x->c, (long long)x->sc);
  AST_CHECK_ARG(x->c / vec <= NT, "ast_instnorm_stats: C=%d too large", x->c);
  if (x->n == 0) return 0;
  const int lanes = x->c / vec, slots = NT / lanes, hw = x->h * x->w;
  const int nblk = stats_blocks(x->n, hw, slots);
  AST_CHECK_ARG((int64_t)x->n * nblk * x->c * 12 <= ast_instnorm_workspace_bytes(x->n, x->c), "ast_instnorm_stats: workspace");
  const int chunk = (hw + nblk - 1) / nblk;
  const size_t smem = (NT + 2 * slots * x->c) * sizeof(float);
  dim3 grid(nblk, x->n);
  cudaStream_t s = (cudaStream_t)stream;
  if (x->dtype == AST_F32) launch_k(in_stats_partial_kernel<float>, grid, NT, smem, s, to_img(x), (float*)workspace, nblk, chunk);
  else launch_k(in_stats_partial_kernel<__nv_bfloat16>, grid, NT, smem, s, to_img(x), (float*)workspace, nblk, chunk);
  launch_k(in_stats_final_kernel, x->n, 128, 0, s, (const float*)workspace, nblk, x->c, hw, eps, mean, rstd);
  count_launch(2);
  count_work(FAM_IN_STATS, 0.0, img_bytes(x));
  AST_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int ast_instnorm_finalize(const double* sums, int32_t n, int32_t c, int32_t hw, float eps, float* mean,
                                     float* rstd, void* stream) {
  AST_CHECK_ARG(sums && mean && rstd && hw > 0, "ast_instnorm_finalize: bad argument");
  const int total = n * c;
  if (total == 0) return 0;
  launch_k(in_finalize_kernel, (total + 255) / 256, 256, 0, (cudaStream_t)stream, sums, total, 1.0 / (double)hw, eps, mean, rstd);
  count_launch();
  count_work(FAM_IN_STATS, 0.0, 0.0);
  AST_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int ast_instnorm_apply(const ast_image* x, const float* mean, const float* rstd, const float* gamma,
                                  const float* beta, const ast_image* residual, const ast_image* out, int32_t pad,
                                  int32_t relu, void* stream) {
  AST_CHECK_ARG(x && mean && rstd && gamma && beta && out, "ast_instnorm_apply: null argument");
  AST_CHECK_ARG(nhwc_ok(x, 4) && nhwc_ok(out, 4), "ast_instnorm_apply: needs NHWC, C %% 4 == 0");
  AST_CHECK_ARG(out->n == x->n && out->c == x->c && out->h == x->h + 2 * pad && out->w == x->w + 2 * pad,
                "ast_instnorm_apply: out must be (h+2p, w+2p)");
  AST_CHECK_ARG(pad < x->h && pad < x->w, "ast_instnorm_apply: pad too large for reflection");
  AST_CHECK_ARG(!residual || (same_shape(residual, x) && residual->sc == 1), "ast_instnorm_apply: residual shape");
  if (x->n == 0) return 0;
  if (int fr = instnorm_apply_staged(x, mean, rstd, gamma, beta, residual, out, pad, relu, (cudaStream_t)stream))
    return fr == 1 ? 0 : fr;
  const long long total = (long long)out->n * out->h * out->w * (x->c / 4);
  const int blocks = (int)min((total + NT - 1) / NT, (long long)num_sms() * 16);
  Img r = residual ? to_img(residual) : null_img();
  cudaStream_t s = (cudaStream_t)stream;
#define LA(TX, TO) launch_k(in_apply_kernel<TX, TO>, blocks, NT, 0, s, to_img(x), mean, rstd, gamma, beta, r, to_img(out), pad, relu)
  if (x->dtype == AST_F32 && out->dtype == AST_F32) LA(float, float);
  else if (x->dtype == AST_BF16 && out->dtype == AST_BF16) LA(__nv_bfloat16, __nv_bfloat16);
  else if (x->dtype == AST_F32 && out->dtype == AST_BF16) LA(float, __nv_bfloat16);
  else LA(__nv_bfloat16, float);
#undef LA
  count_launch();
  count_work(FAM_IN_APPLY, 0.0, img_bytes(x) + img_bytes(out) + img_bytes(residual));
  AST_CUDA_LAUNCH_CHECK();
  return 0;
}

static int bwd_check(const char* who, const ast_image* x, const ast_image* gpad, int pad, const ast_image* gextra) {
  AST_CHECK_ARG(nhwc_ok(x, 4), "%s: needs NHWC, C %% 4 == 0", who);
  AST_CHECK_ARG(x->c / 4 <= NT, "%s: C too large", who);
  AST_CHECK_ARG(!gpad || (gpad->n == x->n && gpad->c == x->c && gpad->h == x->h + 2 * pad && gpad->w == x->w + 2 * pad && gpad->sc == 1),
                "%s: gpad must be (h+2p, w+2p) NHWC", who);
  AST_CHECK_ARG(!gextra || (same_shape(gextra, x) && gextra->sc == 1), "%s: gextra shape", who);
  AST_CHECK_ARG(pad < x->h && pad < x->w, "%s: pad too large", who);
  return 0;
}

extern "C" int ast_instnorm_bwd_stats(const ast_image* x, const float* mean, const float* rstd, const float* gamma,
                                      const float* beta, const ast_image* gpad, int32_t pad, const ast_image* gextra,
                                      int32_t relu, float* s1, float* s2, void* stream) {
  AST_CHECK_ARG(x && mean && rstd && gamma && beta && s1 && s2, "ast_instnorm_bwd_stats: null argument");
  if (int e = bwd_check("ast_instnorm_bwd_stats", x, gpad, pad, gextra)) return e;
  if (x->n == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (!(relu & AST_IN_SUMS_ZEROED)) {
    cudaMemsetAsync(s1, 0, sizeof(float) * x->n * x->c, s);
    cudaMemsetAsync(s2, 0, sizeof(float) * x->n * x->c, s);
  }
  relu &= 1;
  if (int fr = instnorm_bwd_stats_staged(x, mean, rstd, gamma, beta, gpad, pad, gextra, relu, s1, s2, s))
    return fr == 1 ? 0 : fr;
  const int lanes = x->c / 4, slots = NT / lanes, hw = x->h * x->w;
  const int nblk = stats_blocks(x->n, hw, slots);
  const int chunk = (hw + nblk - 1) / nblk;
  const size_t smem = 2 * slots * x->c * sizeof(float);
  dim3 grid(nblk, x->n);
  Img gp = gpad ? to_img(gpad) : null_img(), ge = gextra ? to_img(gextra) : null_img();
  if (x->dtype == AST_F32)
    launch_k(in_bwd_stats_kernel<float>, grid, NT, smem, s, to_img(x), mean, rstd, gamma, beta, gp, pad, ge, relu, s1, s2, chunk);
  else
    launch_k(in_bwd_stats_kernel<__nv_bfloat16>, grid, NT, smem, s, to_img(x), mean, rstd, gamma, beta, gp, pad, ge, relu, s1, s2, chunk);
  count_launch();
  count_work(FAM_IN_BWD, 0.0, 0.0);
  AST_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int ast_instnorm_bwd_apply(const ast_image* x, const float* mean, const float* rstd, const float* gamma,
                                      const float* beta, const ast_image* gpad, int32_t pad, const ast_image* gextra,
                                      int32_t relu, const float* s1, const float* s2, const ast_image* dx,
                                      const ast_image* gtotal, void* stream) {
  AST_CHECK_ARG(x && mean && rstd && gamma && beta && s1 && s2 && dx, "ast_instnorm_bwd_apply: null argument");
  if (int e = bwd_check("ast_instnorm_bwd_apply", x, gpad, pad, gextra)) return e;
  AST_CHECK_ARG(same_shape(dx, x) && dx->sc == 1, "ast_instnorm_bwd_apply: dx shape");
  AST_CHECK_ARG(!gtotal || (same_shape(gtotal, x) && gtotal->sc == 1), "ast_instnorm_bwd_apply: gtotal shape");
  if (x->n == 0) return 0;
  relu &= 1;
  if (int fr = instnorm_bwd_apply_staged(x, mean, rstd, gamma, beta, gpad, pad, gextra, relu, s1, s2, dx, gtotal,
                                       (cudaStream_t)stream))
    return fr == 1 ? 0 : fr;
  const long long total = (long long)x->n * x->h * x->w * (x->c / 4);
  const int blocks = (int)min((total + NT - 1) / NT, (long long)num_sms() * 16);
  Img gp = gpad ? to_img(gpad) : null_img(), ge = gextra ? to_img(gextra) : null_img();
  Img gt = gtotal ? to_img(gtotal) : null_img();
  cudaStream_t s = (cudaStream_t)stream;
#define LB(TX, TO) launch_k(in_bwd_apply_kernel<TX, TO>, blocks, NT, 0, s, to_img(x), mean, rstd, gamma, beta, gp, pad, ge, relu, s1, s2, to_img(dx), gt)
  if (x->dtype == AST_F32 && dx->dtype == AST_F32) LB(float, float);
  else if (x->dtype == AST_BF16 && dx->dtype == AST_BF16) LB(__nv_bfloat16, __nv_bfloat16);
  else if (x->dtype == AST_F32 && dx->dtype == AST_BF16) LB(float, __nv_bfloat16);
  else LB(__nv_bfloat16, float);
#undef LB
  count_launch();
  count_work(FAM_IN_BWD, 0.0, 2.0 * img_bytes(x) + img_bytes(dx) + img_bytes(gtotal));
  AST_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int ast_instnorm_bwd(const ast_image* x, const float* mean, const float* rstd, const float* gamma,
                                const float* beta, const ast_image* gpad, int32_t pad, const ast_image* gextra,
                                int32_t relu, float* s1, float* s2, int32_t* arrive, const ast_image* dx,
                                const ast_image* gtotal, void* stream) {
  AST_CHECK_ARG(x && mean && rstd && gamma && beta && s1 && s2 && dx, "ast_instnorm_bwd: null argument");
  if (int e = bwd_check("ast_instnorm_bwd", x, gpad, pad, gextra)) return e;
  AST_CHECK_ARG(same_shape(dx, x) && dx->sc == 1, "ast_instnorm_bwd: dx shape");
  AST_CHECK_ARG(!gtotal || (same_shape(gtotal, x) && gtotal->sc == 1), "ast_instnorm_bwd: gtotal shape");
  if (x->n == 0) return 0;
  if (arrive && (relu & AST_IN_SUMS_ZEROED)) {
    if (int fr = instnorm_bwd_fused_staged(x, mean, rstd, gamma, beta, gpad, pad, gextra, relu & 1, s1, s2, arrive, dx, gtotal,
                                           (cudaStream_t)stream))
      return fr == 1 ? 0 : fr;
  }
  if (int e = ast_instnorm_bwd_stats(x, mean, rstd, gamma, beta, gpad, pad, gextra, relu, s1, s2, stream)) return e;
  return ast_instnorm_bwd_apply(x, mean, rstd, gamma, beta, gpad, pad, gextra, relu, s1, s2, dx, gtotal, stream);
}
