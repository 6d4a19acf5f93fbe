// Vectorised InstanceNorm apply / backward kernels (the HBM-roofline kernels of the path).
//
// Thread mapping: a thread owns ONE 16-byte channel group (8 bf16 / 4 fp32 channels) for the whole kernel, so the
// per-(n,c) constants live in registers; consecutive threads cover consecutive channel groups of a pixel, then
// consecutive pixels -> every warp access is a run of full 128-byte lines.  To keep the register count low enough
// for 4 blocks/SM (memory-level parallelism is what these kernels live on), the per-channel math is folded into few
// constants:   y  = A*x + D                      (A = gamma*rstd, D = beta - A*mean; same expression fwd and bwd)
//              dx = A*g + B*x + C                (B = -A*rstd*s2/HW, C = -A*s1/HW - B*mean)
//              s1 = sum g',  s2 = rstd*(sum g'*x - mean*sum g')
// One pixel costs one integer division (amortised over 8 channels); reflection folding is only evaluated on the
// border ring.  These replace the generic kernels of norm.cu whenever x / out / dx share one dtype and the tensors
// are NHWC with 16-byte-aligned pixel strides (always true inside the engine).
#include "common.cuh"

namespace ast {

constexpr int NF_THREADS = 256;

template <int VEC>
__device__ __forceinline__ void ldv_img(const Img& im, long long off, float* v) {
  if (im.dtype == AST_F32) {
#pragma unroll
    for (int e = 0; e < VEC; e += 4) ld4((const float*)im.ptr + off + e, v + e);
  } else {
    if (VEC == 8) Vec16<__nv_bfloat16>::load((const __nv_bfloat16*)im.ptr + off, v);
    else ld4((const __nv_bfloat16*)im.ptr + off, v);
  }
}
template <int VEC>
__device__ __forceinline__ void stv_img(const Img& im, long long off, const float* v) {
  if (im.dtype == AST_F32) {
#pragma unroll
    for (int e = 0; e < VEC; e += 4) st4((float*)im.ptr + off + e, v + e);
  } else {
    if (VEC == 8) Vec16<__nv_bfloat16>::store((__nv_bfloat16*)im.ptr + off, v);
    else st4((__nv_bfloat16*)im.ptr + off, v);
  }
}

template <typename T> struct Raw16;
template <> struct Raw16<float> {
  __device__ static __forceinline__ void unpack(const uint4& r, float* v) {
    v[0] = __uint_as_float(r.x); v[1] = __uint_as_float(r.y); v[2] = __uint_as_float(r.z); v[3] = __uint_as_float(r.w);
  }
  __device__ static __forceinline__ uint4 pack(const float* v) {
    return make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
  }
};
template <> struct Raw16<__nv_bfloat16> {
  __device__ static __forceinline__ void unpack(const uint4& r, float* v) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
  }
  __device__ static __forceinline__ uint4 pack(const float* v) {
    uint4 r;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    return r;
  }
};
template <typename T>
__device__ __forceinline__ uint4 ldraw(const Img& im, long long off) {
  return *reinterpret_cast<const uint4*>((const T*)im.ptr + off);
}

constexpr int UNR = 4;   // pixels in flight per thread: all 16-byte loads of a batch are issued before any is used

// ---------------------------------------------------------------- forward apply
template <typename T>
__global__ void __launch_bounds__(NF_THREADS, 2)
in_apply_fast_kernel(Img x, const float* __restrict__ mean, const float* __restrict__ rstd,
                     const float* __restrict__ gamma, const float* __restrict__ beta, Img res, Img out, int pad,
                     int relu, int chunk) {
  constexpr int VEC = Vec16<T>::N;
  const int C = x.c, lanes = C / VEC, slots = NF_THREADS / lanes;
  const int lane = threadIdx.x % lanes, slot = threadIdx.x / lanes;
  const int n = blockIdx.y, c = lane * VEC;
  float A[VEC], D[VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) {
    A[e] = gamma[c + e] * rstd[n * C + c + e];
    D[e] = beta[c + e] - A[e] * mean[n * C + c + e];
  }
  const int npix = out.h * out.w;
  const int pbeg = blockIdx.x * chunk, pend = min(npix, pbeg + chunk);
  for (int p0 = pbeg + slot; p0 < pend; p0 += UNR * slots) {
    uint4 xr[UNR], rr[UNR];
    long long oo[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int p = p0 + u * slots;
      oo[u] = -1;
      if (p < pend) {
        const int oy = p / out.w, ox = p - oy * out.w;
        const int i = reflect_idx(oy - pad, x.h), j = reflect_idx(ox - pad, x.w);
        xr[u] = ldraw<T>(x, img_off(x, n, i, j, c));
        if (res.ptr) rr[u] = ldraw<T>(res, img_off(res, n, i, j, c));
        oo[u] = img_off(out, n, oy, ox, c);
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      if (oo[u] < 0) continue;
      float v[VEC];
      Raw16<T>::unpack(xr[u], v);
#pragma unroll
      for (int e = 0; e < VEC; ++e) v[e] = fmaf(A[e], v[e], D[e]);
      if (res.ptr) {
        float r[VEC];
        Raw16<T>::unpack(rr[u], r);
#pragma unroll
        for (int e = 0; e < VEC; ++e) v[e] += r[e];
      }
      if (relu) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) v[e] = fmaxf(v[e], 0.f);
      }
      *reinterpret_cast<uint4*>((T*)out.ptr + oo[u]) = Raw16<T>::pack(v);
    }
  }
}

// ---------------------------------------------------------------- backward
// g' = (fold_reflect(gpad) + gextra) * relu_mask.  The centre gpad value (and gextra) arrive pre-loaded in `g`;
// only pixels on the border ring read the (up to 8) mirrored positions here.
template <typename T, int VEC>
__device__ __forceinline__ void fold_border(const Img& gpad, int pad, int H, int W, int n, int i, int j, int c, float* g) {
  const bool by = (i <= pad || i >= H - 1 - pad), bx = (j <= pad || j >= W - 1 - pad);
  if (!(by || bx)) return;
  int rr[3], cc[3], nr = 1, nc = 1;
  rr[0] = i + pad; cc[0] = j + pad;
  if (i >= 1 && i <= pad) rr[nr++] = pad - i;
  if (i <= H - 2 && i >= H - 1 - pad) rr[nr++] = pad + 2 * (H - 1) - i;
  if (j >= 1 && j <= pad) cc[nc++] = pad - j;
  if (j <= W - 2 && j >= W - 1 - pad) cc[nc++] = pad + 2 * (W - 1) - j;
  for (int a_ = 0; a_ < nr; ++a_)
    for (int b_ = 0; b_ < nc; ++b_) {
      if (a_ == 0 && b_ == 0) continue;
      float t[VEC];
      Raw16<T>::unpack(ldraw<T>(gpad, img_off(gpad, n, rr[a_], cc[b_], c)), t);
#pragma unroll
      for (int e = 0; e < VEC; ++e) g[e] += t[e];
    }
}

// loads of one batch: x, centre gpad, gextra (raw 16-byte vectors), pixel coordinates
#define IN_BWD_LOAD_BATCH()                                                                   \
  uint4 xr[UNR], gr[UNR], er[UNR];                                                            \
  int pi[UNR], pj[UNR];                                                                       \
  _Pragma("unroll") for (int u = 0; u < UNR; ++u) {                                           \
    const int p = p0 + u * slots;                                                             \
    pi[u] = -1;                                                                               \
    if (p < pend) {                                                                           \
      const int i = p / x.w, j = p - i * x.w;                                                 \
      pi[u] = i; pj[u] = j;                                                                   \
      xr[u] = ldraw<T>(x, img_off(x, n, i, j, c));                                            \
      if (gpad.ptr) gr[u] = ldraw<T>(gpad, img_off(gpad, n, i + pad, j + pad, c));            \
      if (gextra.ptr) er[u] = ldraw<T>(gextra, img_off(gextra, n, i, j, c));                  \
    }                                                                                         \
  }

#define IN_BWD_GPRIME(u)                                                                      \
  float xv[VEC], g[VEC];                                                                      \
  Raw16<T>::unpack(xr[u], xv);                                                                \
  if (gpad.ptr) {                                                                             \
    Raw16<T>::unpack(gr[u], g);                                                               \
    if (pad > 0) fold_border<T, VEC>(gpad, pad, x.h, x.w, n, pi[u], pj[u], c, g);             \
  } else {                                                                                    \
    _Pragma("unroll") for (int e = 0; e < VEC; ++e) g[e] = 0.f;                               \
  }                                                                                           \
  if (gextra.ptr) {                                                                           \
    float t[VEC];                                                                             \
    Raw16<T>::unpack(er[u], t);                                                               \
    _Pragma("unroll") for (int e = 0; e < VEC; ++e) g[e] += t[e];                             \
  }                                                                                           \
  _Pragma("unroll") for (int e = 0; e < VEC; ++e)                                             \
    if (relu && !(fmaf(A[e], xv[e], D[e]) > 0.f)) g[e] = 0.f;   /* same expression as the forward apply */

template <typename T>
__global__ void __launch_bounds__(NF_THREADS, 2)
in_bwd_stats_fast_kernel(Img x, const float* __restrict__ mean, const float* __restrict__ rstd,
                         const float* __restrict__ gamma, const float* __restrict__ beta, Img gpad, int pad, Img gextra,
                         int relu, float* __restrict__ s1o, float* __restrict__ s2o, int chunk) {
  constexpr int VEC = Vec16<T>::N;
  extern __shared__ float sm[];
  const int C = x.c, lanes = C / VEC, slots = NF_THREADS / lanes;
  const int lane = threadIdx.x % lanes, slot = threadIdx.x / lanes;
  const int n = blockIdx.y, c = lane * VEC;
  float A[VEC], D[VEC], t1[VEC], t2[VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) {
    A[e] = gamma[c + e] * rstd[n * C + c + e];
    D[e] = beta[c + e] - A[e] * mean[n * C + c + e];
    t1[e] = 0.f; t2[e] = 0.f;
  }
  const int npix = x.h * x.w;
  const int pbeg = blockIdx.x * chunk, pend = min(npix, pbeg + chunk);
  for (int p0 = pbeg + slot; p0 < pend; p0 += UNR * slots) {
    IN_BWD_LOAD_BATCH()
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      if (pi[u] < 0) continue;
      IN_BWD_GPRIME(u)
#pragma unroll
      for (int e = 0; e < VEC; ++e) { t1[e] += g[e]; t2[e] = fmaf(g[e], xv[e], t2[e]); }
    }
  }
  float* r1 = sm;
  float* r2 = sm + slots * C;
#pragma unroll
  for (int e = 0; e < VEC; ++e) { r1[slot * C + c + e] = t1[e]; r2[slot * C + c + e] = t2[e]; }
  __syncthreads();
  for (int cc = threadIdx.x; cc < C; cc += NF_THREADS) {
    float u1 = 0.f, u2 = 0.f;
    for (int s = 0; s < slots; ++s) { u1 += r1[s * C + cc]; u2 += r2[s * C + cc]; }
    // s2 = sum g*xhat = rstd * (sum g*x - mean * sum g)
    atomicAdd(s1o + n * C + cc, u1);
    atomicAdd(s2o + n * C + cc, rstd[n * C + cc] * (u2 - mean[n * C + cc] * u1));
  }
}

template <typename T>
__global__ void __launch_bounds__(NF_THREADS, 2)
in_bwd_apply_fast_kernel(Img x, const float* __restrict__ mean, const float* __restrict__ rstd,
                         const float* __restrict__ gamma, const float* __restrict__ beta, Img gpad, int pad, Img gextra,
                         int relu, const float* __restrict__ s1, const float* __restrict__ s2, Img dx, Img gtotal,
                         int chunk) {
  constexpr int VEC = Vec16<T>::N;
  const int C = x.c, lanes = C / VEC, slots = NF_THREADS / lanes;
  const int lane = threadIdx.x % lanes, slot = threadIdx.x / lanes;
  const int n = blockIdx.y, c = lane * VEC;
  const float inv_hw = 1.f / (float)(x.h * x.w);
  float A[VEC], B[VEC], Cc[VEC], D[VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) {
    const float mu = mean[n * C + c + e], rs = rstd[n * C + c + e];
    A[e] = gamma[c + e] * rs;
    D[e] = beta[c + e] - A[e] * mu;
    B[e] = -A[e] * rs * s2[n * C + c + e] * inv_hw;
    Cc[e] = -A[e] * s1[n * C + c + e] * inv_hw - B[e] * mu;
  }
  const int npix = x.h * x.w;
  const int pbeg = blockIdx.x * chunk, pend = min(npix, pbeg + chunk);
  for (int p0 = pbeg + slot; p0 < pend; p0 += UNR * slots) {
    IN_BWD_LOAD_BATCH()
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      if (pi[u] < 0) continue;
      IN_BWD_GPRIME(u)
      if (gtotal.ptr) *reinterpret_cast<uint4*>((T*)gtotal.ptr + img_off(gtotal, n, pi[u], pj[u], c)) = Raw16<T>::pack(g);
#pragma unroll
      for (int e = 0; e < VEC; ++e) xv[e] = fmaf(A[e], g[e], fmaf(B[e], xv[e], Cc[e]));
      *reinterpret_cast<uint4*>((T*)dx.ptr + img_off(dx, n, pi[u], pj[u], c)) = Raw16<T>::pack(xv);
    }
  }
}

static bool fast_ok(const ast_image* im, int vec) {
  return im->sc == 1 && im->c % vec == 0 && im->sw % vec == 0 && im->sh % vec == 0 && im->sn % vec == 0 &&
         ((uintptr_t)im->ptr & 15) == 0;
}

static int grid_chunks(int n, int npix, int slots, int* chunk) {
  int nblk = (8 * num_sms() + n - 1) / n;
  const int maxb = (npix + 2 * UNR * slots - 1) / (2 * UNR * slots);   // at least two batches of UNR pixels per thread
  if (nblk > maxb) nblk = maxb;
  if (nblk < 1) nblk = 1;
  *chunk = (npix + nblk - 1) / nblk;
  return (npix + *chunk - 1) / *chunk;
}

// Returns 1 if the fast kernel was launched, 0 if the shapes/dtypes need the generic kernel, <0 / >0 on error.
int instnorm_apply_fast(const ast_image* x, const float* mean, const float* rstd, const float* gamma, const float* beta,
                        const ast_image* residual, const ast_image* out, int pad, int relu, cudaStream_t s) {
  const int vec = x->dtype == AST_F32 ? 4 : 8;
  if (x->dtype != out->dtype || !fast_ok(x, vec) || !fast_ok(out, vec) || NF_THREADS % (x->c / vec) != 0) return 0;
  if (residual && (!fast_ok(residual, vec) || residual->dtype != x->dtype)) return 0;
  const int slots = NF_THREADS / (x->c / vec);
  int chunk;
  const int nblk = grid_chunks(x->n, out->h * out->w, slots, &chunk);
  dim3 grid(nblk, x->n);
  Img r = residual ? to_img(residual) : null_img();
  if (x->dtype == AST_F32)
    in_apply_fast_kernel<float><<<grid, NF_THREADS, 0, s>>>(to_img(x), mean, rstd, gamma, beta, r, to_img(out), pad, relu, chunk);
  else
    in_apply_fast_kernel<__nv_bfloat16><<<grid, NF_THREADS, 0, s>>>(to_img(x), mean, rstd, gamma, beta, r, to_img(out), pad, relu, chunk);
  count_launch();
  AST_CUDA_LAUNCH_CHECK();
  return 1;
}

int instnorm_bwd_stats_fast(const ast_image* x, const float* mean, const float* rstd, const float* gamma,
                            const float* beta, const ast_image* gpad, int pad, const ast_image* gextra, int relu,
                            float* s1, float* s2, cudaStream_t s) {
  const int vec = x->dtype == AST_F32 ? 4 : 8;
  if (!fast_ok(x, vec) || NF_THREADS % (x->c / vec) != 0) return 0;
  if ((gpad && (!fast_ok(gpad, vec) || gpad->dtype != x->dtype)) || (gextra && (!fast_ok(gextra, vec) || gextra->dtype != x->dtype))) return 0;
  const int slots = NF_THREADS / (x->c / vec);
  int chunk;
  const int nblk = grid_chunks(x->n, x->h * x->w, slots, &chunk);
  dim3 grid(nblk, x->n);
  const size_t smem = 2 * (size_t)slots * x->c * sizeof(float);
  Img gp = gpad ? to_img(gpad) : null_img(), ge = gextra ? to_img(gextra) : null_img();
  if (x->dtype == AST_F32)
    in_bwd_stats_fast_kernel<float><<<grid, NF_THREADS, smem, s>>>(to_img(x), mean, rstd, gamma, beta, gp, pad, ge, relu, s1, s2, chunk);
  else
    in_bwd_stats_fast_kernel<__nv_bfloat16><<<grid, NF_THREADS, smem, s>>>(to_img(x), mean, rstd, gamma, beta, gp, pad, ge, relu, s1, s2, chunk);
  count_launch();
  AST_CUDA_LAUNCH_CHECK();
  return 1;
}

int instnorm_bwd_apply_fast(const ast_image* x, const float* mean, const float* rstd, const float* gamma,
                            const float* beta, const ast_image* gpad, int pad, const ast_image* gextra, int relu,
                            const float* s1, const float* s2, const ast_image* dx, const ast_image* gtotal,
                            cudaStream_t s) {
  const int vec = x->dtype == AST_F32 ? 4 : 8;
  if (x->dtype != dx->dtype || !fast_ok(x, vec) || !fast_ok(dx, vec) || NF_THREADS % (x->c / vec) != 0) return 0;
  if ((gpad && (!fast_ok(gpad, vec) || gpad->dtype != x->dtype)) || (gextra && (!fast_ok(gextra, vec) || gextra->dtype != x->dtype)) ||
      (gtotal && (!fast_ok(gtotal, vec) || gtotal->dtype != x->dtype))) return 0;
  const int slots = NF_THREADS / (x->c / vec);
  int chunk;
  const int nblk = grid_chunks(x->n, x->h * x->w, slots, &chunk);
  dim3 grid(nblk, x->n);
  Img gp = gpad ? to_img(gpad) : null_img(), ge = gextra ? to_img(gextra) : null_img();
  Img gt = gtotal ? to_img(gtotal) : null_img();
  if (x->dtype == AST_F32)
    in_bwd_apply_fast_kernel<float><<<grid, NF_THREADS, 0, s>>>(to_img(x), mean, rstd, gamma, beta, gp, pad, ge, relu, s1, s2, to_img(dx), gt, chunk);
  else
    in_bwd_apply_fast_kernel<__nv_bfloat16><<<grid, NF_THREADS, 0, s>>>(to_img(x), mean, rstd, gamma, beta, gp, pad, ge, relu, s1, s2, to_img(dx), gt, chunk);
  count_launch();
  AST_CUDA_LAUNCH_CHECK();
  return 1;
}

}  // namespace ast
