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

#ifndef NF_UNR
#define NF_UNR 4
#endif
#ifndef NF_MINB
#define NF_MINB 2
#endif
#ifndef NF_LDG
#define NF_LDG 1
#endif
#ifndef NF_FOLD_INLINE
#define NF_FOLD_INLINE 1
#endif
constexpr int UNR = NF_UNR;   // pixels in flight per thread: all 16-byte loads of a batch are issued before any is used

// VEC consecutive per-channel floats (16-byte aligned: VEC is 4 or 8 and the offset a multiple of VEC)
template <int VEC>
__device__ __forceinline__ void ldc(const float* __restrict__ p, float* v) {
#pragma unroll
  for (int e = 0; e < VEC; e += 4) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p + e));
    v[e] = t.x; v[e + 1] = t.y; v[e + 2] = t.z; v[e + 3] = t.w;
  }
}

// Lean addressing: 32-bit element offsets (the host checks that every tensor spans < 2^31 elements) and a multiply-high
// instead of the per-pixel integer division (magic = ceil(2^32 / W), exact for all p < H*W - checked on the host; 0
// selects the plain division).  The first version of these kernels spent most of its issue slots on 64-bit index
// arithmetic (ncu: issue-active 46 % at 23 % occupancy, 128 registers); see profiles/r01_summary.md.
struct Lin {
  const char* ptr;
  int sn, sh, sw;
};
template <typename T>
__device__ __forceinline__ uint4 ldlin(const Lin& t, int off) {
#if NF_LDG
  return __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(t.ptr) + off));
#else
  return *reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(t.ptr) + off);
#endif
}
template <typename T>
__device__ __forceinline__ void stlin(const Lin& t, int off, const uint4& v) {
  *reinterpret_cast<uint4*>(reinterpret_cast<T*>(const_cast<char*>(t.ptr)) + off) = v;
}
__device__ __forceinline__ void pix_decode(int p, int W, unsigned magic, int& i, int& j) {
  i = magic ? (int)__umulhi((unsigned)p, magic) : p / W;
  j = p - i * W;
}

struct NfShape {
  int C, H, W;          // logical (unpadded) image
  int pad;
  int rows;             // image rows per block
};

// Loop structure shared by the three kernels: a block owns `rows` consecutive image rows of one image; within a row a
// thread visits the pixels slot, slot + slots, ... in batches of UNR (for the transform net W*C = 8192 elements, so
// one batch is exactly one row).  The row index is a loop variable: no per-pixel division, row offsets and the
// border-row test are hoisted, a pixel costs one multiply-add per tensor.

// ---------------------------------------------------------------- forward apply
template <typename T>
__global__ void __launch_bounds__(NF_THREADS, NF_MINB)
in_apply_fast_kernel(Lin x, const float* __restrict__ mean, const float* __restrict__ rstd,
                     const float* __restrict__ gamma, const float* __restrict__ beta, Lin res, Lin out, NfShape sh,
                     int relu) {
  pdl_sync();   // programmatic dependent launch: see common.cuh
  constexpr int VEC = Vec16<T>::N;
  const int C = sh.C, lanes = C / VEC, slots = NF_THREADS / lanes;
  const int lane = threadIdx.x % lanes, slot = threadIdx.x / lanes;
  const int n = blockIdx.y, c = lane * VEC;
  float A[VEC], D[VEC];
  {
    float mu[VEC];
    ldc<VEC>(gamma + c, A); ldc<VEC>(rstd + n * C + c, D); ldc<VEC>(mean + n * C + c, mu);
#pragma unroll
    for (int e = 0; e < VEC; ++e) A[e] *= D[e];
    ldc<VEC>(beta + c, D);
#pragma unroll
    for (int e = 0; e < VEC; ++e) D[e] -= A[e] * mu[e];
  }
  const int OH = sh.H + 2 * sh.pad, OW = sh.W + 2 * sh.pad;
  const int rbeg = blockIdx.x * sh.rows, rend = min(OH, rbeg + sh.rows);
  const int xb = n * x.sn + c, rb = n * res.sn + c, ob = n * out.sn + c;
  for (int oy = rbeg; oy < rend; ++oy) {
    const int i = reflect_idx(oy - sh.pad, sh.H);
    const int xrow = xb + i * x.sh, rrow = rb + i * res.sh, orow = ob + oy * out.sh;
    for (int ox0 = slot; ox0 < OW; ox0 += UNR * slots) {
      uint4 xr[UNR], rr[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int ox = ox0 + u * slots;
        if (ox < OW) {
          const int j = reflect_idx(ox - sh.pad, sh.W);
          xr[u] = ldlin<T>(x, xrow + j * x.sw);
          if (res.ptr) rr[u] = ldlin<T>(res, rrow + j * res.sw);
        }
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int ox = ox0 + u * slots;
        if (ox >= OW) continue;
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
        stlin<T>(out, orow + ox * out.sw, Raw16<T>::pack(v));
      }
    }
  }
}

// ---------------------------------------------------------------- backward
// g' = (fold_reflect(gpad) + gextra) * relu_mask.  The centre gpad value (and gextra) arrive pre-loaded in `g`;
// only pixels on the border ring read the (up to 8) mirrored positions here.
template <typename T, int VEC>
__device__ __forceinline__ void fold_border(const Lin& gpad, int gbase, int pad, int H, int W, int i, int j, float* g) {
  // mirrored rows / columns of the padded gradient that fold onto (i, j); -1 = none
  const int r1 = (i >= 1 && i <= pad) ? pad - i : -1;
  const int r2 = (i <= H - 2 && i >= H - 1 - pad) ? pad + 2 * (H - 1) - i : -1;
  const int c1 = (j >= 1 && j <= pad) ? pad - j : -1;
  const int c2 = (j <= W - 2 && j >= W - 1 - pad) ? pad + 2 * (W - 1) - j : -1;
#pragma unroll
  for (int a_ = 0; a_ < 3; ++a_) {
    const int r = a_ == 0 ? i + pad : (a_ == 1 ? r1 : r2);
    if (r < 0) continue;
#pragma unroll
    for (int b_ = 0; b_ < 3; ++b_) {
      const int cc = b_ == 0 ? j + pad : (b_ == 1 ? c1 : c2);
      if (cc < 0 || (a_ == 0 && b_ == 0)) continue;
      float t[VEC];
      Raw16<T>::unpack(ldlin<T>(gpad, gbase + r * gpad.sh + cc * gpad.sw), t);
#pragma unroll
      for (int e = 0; e < VEC; ++e) g[e] += t[e];
    }
  }
}

// per-row set-up, then the loads of one batch: x, centre gpad, gextra (raw 16-byte vectors)
#define IN_BWD_ROW_SETUP()                                                                    \
  const bool brow = sh.pad > 0 && (i <= sh.pad || i >= sh.H - 1 - sh.pad);                    \
  const int xrow = xb + i * x.sh, grow = gb + (i + sh.pad) * gpad.sh + sh.pad * gpad.sw,      \
            erow = eb + i * gextra.sh;

#define IN_BWD_LOAD_BATCH()                                                                   \
  uint4 xr[UNR], gr[UNR], er[UNR];                                                            \
  _Pragma("unroll") for (int u = 0; u < UNR; ++u) {                                           \
    const int j = j0 + u * slots;                                                             \
    if (j < sh.W) {                                                                           \
      xr[u] = ldlin<T>(x, xrow + j * x.sw);                                                   \
      if (gpad.ptr) gr[u] = ldlin<T>(gpad, grow + j * gpad.sw);                               \
      if (gextra.ptr) er[u] = ldlin<T>(gextra, erow + j * gextra.sw);                         \
    }                                                                                         \
  }

#define IN_BWD_GPRIME(u)                                                                      \
  float xv[VEC], g[VEC];                                                                      \
  Raw16<T>::unpack(xr[u], xv);                                                                \
  if (gpad.ptr) {                                                                             \
    Raw16<T>::unpack(gr[u], g);                                                               \
    if (sh.pad > 0 && (brow || j <= sh.pad || j >= sh.W - 1 - sh.pad))                        \
      fold_border<T, VEC>(gpad, gb, sh.pad, sh.H, sh.W, i, j, g);                             \
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
__global__ void __launch_bounds__(NF_THREADS, NF_MINB)
in_bwd_stats_fast_kernel(Lin x, const float* __restrict__ mean, const float* __restrict__ rstd,
                         const float* __restrict__ gamma, const float* __restrict__ beta, Lin gpad, Lin gextra,
                         NfShape sh, int relu, float* __restrict__ s1o, float* __restrict__ s2o) {
  pdl_sync();   // programmatic dependent launch: see common.cuh
  constexpr int VEC = Vec16<T>::N;
  extern __shared__ float sm[];
  const int C = sh.C, lanes = C / VEC, slots = NF_THREADS / lanes;
  const int lane = threadIdx.x % lanes, slot = threadIdx.x / lanes;
  const int n = blockIdx.y, c = lane * VEC;
  float A[VEC], D[VEC], t1[VEC], t2[VEC];
  {
    float mu[VEC];
    ldc<VEC>(gamma + c, A); ldc<VEC>(rstd + n * C + c, D); ldc<VEC>(mean + n * C + c, mu);
#pragma unroll
    for (int e = 0; e < VEC; ++e) A[e] *= D[e];
    ldc<VEC>(beta + c, D);
#pragma unroll
    for (int e = 0; e < VEC; ++e) { D[e] -= A[e] * mu[e]; t1[e] = 0.f; t2[e] = 0.f; }
  }
  const int rbeg = blockIdx.x * sh.rows, rend = min(sh.H, rbeg + sh.rows);
  const int xb = n * x.sn + c, gb = n * gpad.sn + c, eb = n * gextra.sn + c;
  for (int i = rbeg; i < rend; ++i) {
    IN_BWD_ROW_SETUP()
    for (int j0 = slot; j0 < sh.W; j0 += UNR * slots) {
      IN_BWD_LOAD_BATCH()
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int j = j0 + u * slots;
        if (j >= sh.W) continue;
        IN_BWD_GPRIME(u)
#pragma unroll
        for (int e = 0; e < VEC; ++e) { t1[e] += g[e]; t2[e] = fmaf(g[e], xv[e], t2[e]); }
      }
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
__global__ void __launch_bounds__(NF_THREADS, NF_MINB)
in_bwd_apply_fast_kernel(Lin x, const float* __restrict__ mean, const float* __restrict__ rstd,
                         const float* __restrict__ gamma, const float* __restrict__ beta, Lin gpad, Lin gextra,
                         NfShape sh, int relu, const float* __restrict__ s1, const float* __restrict__ s2, Lin dx,
                         Lin gtotal) {
  pdl_sync();   // programmatic dependent launch: see common.cuh
  constexpr int VEC = Vec16<T>::N;
  const int C = sh.C, lanes = C / VEC, slots = NF_THREADS / lanes;
  const int lane = threadIdx.x % lanes, slot = threadIdx.x / lanes;
  const int n = blockIdx.y, c = lane * VEC;
  const float inv_hw = 1.f / (float)(sh.H * sh.W);
  float A[VEC], B[VEC], Cc[VEC], D[VEC];
  {
    float mu[VEC], rs[VEC];
    ldc<VEC>(gamma + c, A); ldc<VEC>(rstd + n * C + c, rs); ldc<VEC>(mean + n * C + c, mu);
    ldc<VEC>(beta + c, D); ldc<VEC>(s2 + n * C + c, B); ldc<VEC>(s1 + n * C + c, Cc);
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      A[e] *= rs[e];
      D[e] -= A[e] * mu[e];
      B[e] = -A[e] * rs[e] * B[e] * inv_hw;
      Cc[e] = -A[e] * Cc[e] * inv_hw - B[e] * mu[e];
    }
  }
  const int rbeg = blockIdx.x * sh.rows, rend = min(sh.H, rbeg + sh.rows);
  const int xb = n * x.sn + c, gb = n * gpad.sn + c, eb = n * gextra.sn + c, db = n * dx.sn + c, tb = n * gtotal.sn + c;
  for (int i = rbeg; i < rend; ++i) {
    IN_BWD_ROW_SETUP()
    const int drow = db + i * dx.sh, trow = tb + i * gtotal.sh;
    for (int j0 = slot; j0 < sh.W; j0 += UNR * slots) {
      IN_BWD_LOAD_BATCH()
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int j = j0 + u * slots;
        if (j >= sh.W) continue;
        IN_BWD_GPRIME(u)
        if (gtotal.ptr) stlin<T>(gtotal, trow + j * gtotal.sw, Raw16<T>::pack(g));
#pragma unroll
        for (int e = 0; e < VEC; ++e) xv[e] = fmaf(A[e], g[e], fmaf(B[e], xv[e], Cc[e]));
        stlin<T>(dx, drow + j * dx.sw, Raw16<T>::pack(xv));
      }
    }
  }
}

static Lin to_lin(const ast_image* im) {
  Lin l;
  if (!im) { l.ptr = nullptr; l.sn = l.sh = l.sw = 0; return l; }
  l.ptr = (const char*)im->ptr; l.sn = (int)im->sn; l.sh = (int)im->sh; l.sw = (int)im->sw;
  return l;
}
// whole tensor addressable with 32-bit element offsets
static bool lin_ok(const ast_image* im) {
  if (!im) return true;
  const long long span = (long long)(im->n - 1) * im->sn + (long long)(im->h - 1) * im->sh + (long long)(im->w - 1) * im->sw + im->c;
  return im->sn >= 0 && im->sh >= 0 && im->sw >= 0 && span < (1ll << 31);
}

static bool fast_ok(const ast_image* im, int vec) {
  return im->sc == 1 && im->c % vec == 0 && im->sw % vec == 0 && im->sh % vec == 0 && im->sn % vec == 0 &&
         ((uintptr_t)im->ptr & 15) == 0;
}

// blocks per image and rows per block: about one wave of resident blocks (2 per SM at <= 128 registers); long blocks
// amortise the per-(n,c) constant set-up
static int grid_rows(int n, int rows_total, int* rows) {
  static const int per_sm = [] { const char* e = getenv("AST_IN_BLOCKS_PER_SM"); return e && atoi(e) > 0 ? atoi(e) : 2; }();
  int nblk = (per_sm * num_sms()) / n;
  if (nblk > rows_total) nblk = rows_total;
  if (nblk < 1) nblk = 1;
  *rows = (rows_total + nblk - 1) / nblk;
  return (rows_total + *rows - 1) / *rows;
}

// norm_staged.cu: bulk-copy staged versions (tried first; 0 = shape not suited, use the register kernels below)
int instnorm_apply_staged(const ast_image* x, const float* mean, const float* rstd, const float* gamma, const float* beta,
                          const ast_image* residual, const ast_image* out, int pad, int relu, cudaStream_t s);
int instnorm_bwd_stats_staged(const ast_image* x, const float* mean, const float* rstd, const float* gamma,
                              const float* beta, const ast_image* gpad, int pad, const ast_image* gextra, int relu,
                              float* s1, float* s2, cudaStream_t s);
int instnorm_bwd_apply_staged(const ast_image* x, const float* mean, const float* rstd, const float* gamma,
                              const float* beta, const ast_image* gpad, int pad, const ast_image* gextra, int relu,
                              const float* s1, const float* s2, const ast_image* dx, const ast_image* gtotal,
                              cudaStream_t s);

// Returns 1 if the fast kernel was launched, 0 if the shapes/dtypes need the generic kernel, <0 / >0 on error.
int instnorm_apply_fast(const ast_image* x, const float* mean, const float* rstd, const float* gamma, const float* beta,
                        const ast_image* residual, const ast_image* out, int pad, int relu, cudaStream_t s) {
  if (int r = instnorm_apply_staged(x, mean, rstd, gamma, beta, residual, out, pad, relu, s)) return r;
  const int vec = x->dtype == AST_F32 ? 4 : 8;
  if (x->dtype != out->dtype || !fast_ok(x, vec) || !fast_ok(out, vec) || NF_THREADS % (x->c / vec) != 0) return 0;
  if (residual && (!fast_ok(residual, vec) || residual->dtype != x->dtype)) return 0;
  if (!lin_ok(x) || !lin_ok(out) || !lin_ok(residual)) return 0;
  NfShape sh;
  sh.C = x->c; sh.H = x->h; sh.W = x->w; sh.pad = pad;
  const int nblk = grid_rows(x->n, out->h, &sh.rows);
  dim3 grid(nblk, x->n);
  if (x->dtype == AST_F32)
    launch_k(in_apply_fast_kernel<float>, grid, NF_THREADS, 0, s, to_lin(x), mean, rstd, gamma, beta, to_lin(residual), to_lin(out), sh, relu);
  else
    launch_k(in_apply_fast_kernel<__nv_bfloat16>, grid, NF_THREADS, 0, s, to_lin(x), mean, rstd, gamma, beta, to_lin(residual), to_lin(out), sh, relu);
  count_launch();
  AST_CUDA_LAUNCH_CHECK();
  return 1;
}

static bool bwd_fast_ok(const ast_image* x, const ast_image* gpad, const ast_image* gextra, int vec) {
  if (!fast_ok(x, vec) || NF_THREADS % (x->c / vec) != 0 || !lin_ok(x)) return false;
  if (gpad && (!fast_ok(gpad, vec) || gpad->dtype != x->dtype || !lin_ok(gpad))) return false;
  if (gextra && (!fast_ok(gextra, vec) || gextra->dtype != x->dtype || !lin_ok(gextra))) return false;
  return true;
}

int instnorm_bwd_stats_fast(const ast_image* x, const float* mean, const float* rstd, const float* gamma,
                            const float* beta, const ast_image* gpad, int pad, const ast_image* gextra, int relu,
                            float* s1, float* s2, cudaStream_t s) {
  if (int r = instnorm_bwd_stats_staged(x, mean, rstd, gamma, beta, gpad, pad, gextra, relu, s1, s2, s)) return r;
  const int vec = x->dtype == AST_F32 ? 4 : 8;
  if (!bwd_fast_ok(x, gpad, gextra, vec)) return 0;
  NfShape sh;
  sh.C = x->c; sh.H = x->h; sh.W = x->w; sh.pad = pad;
  const int nblk = grid_rows(x->n, x->h, &sh.rows);
  dim3 grid(nblk, x->n);
  const size_t smem = 2 * (size_t)(NF_THREADS / (x->c / vec)) * x->c * sizeof(float);
  if (x->dtype == AST_F32)
    launch_k(in_bwd_stats_fast_kernel<float>, grid, NF_THREADS, smem, s, to_lin(x), mean, rstd, gamma, beta, to_lin(gpad), to_lin(gextra), sh, relu, s1, s2);
  else
    launch_k(in_bwd_stats_fast_kernel<__nv_bfloat16>, grid, NF_THREADS, smem, s, to_lin(x), mean, rstd, gamma, beta, to_lin(gpad), to_lin(gextra), sh, relu, s1, s2);
  count_launch();
  AST_CUDA_LAUNCH_CHECK();
  return 1;
}

int instnorm_bwd_apply_fast(const ast_image* x, const float* mean, const float* rstd, const float* gamma,
                            const float* beta, const ast_image* gpad, int pad, const ast_image* gextra, int relu,
                            const float* s1, const float* s2, const ast_image* dx, const ast_image* gtotal,
                            cudaStream_t s) {
  if (int r = instnorm_bwd_apply_staged(x, mean, rstd, gamma, beta, gpad, pad, gextra, relu, s1, s2, dx, gtotal, s)) return r;
  const int vec = x->dtype == AST_F32 ? 4 : 8;
  if (x->dtype != dx->dtype || !bwd_fast_ok(x, gpad, gextra, vec) || !fast_ok(dx, vec) || !lin_ok(dx)) return 0;
  if (gtotal && (!fast_ok(gtotal, vec) || gtotal->dtype != x->dtype || !lin_ok(gtotal))) return 0;
  NfShape sh;
  sh.C = x->c; sh.H = x->h; sh.W = x->w; sh.pad = pad;
  const int nblk = grid_rows(x->n, x->h, &sh.rows);
  dim3 grid(nblk, x->n);
  if (x->dtype == AST_F32)
    launch_k(in_bwd_apply_fast_kernel<float>, grid, NF_THREADS, 0, s, to_lin(x), mean, rstd, gamma, beta, to_lin(gpad), to_lin(gextra), sh, relu, s1, s2, to_lin(dx), to_lin(gtotal));
  else
    launch_k(in_bwd_apply_fast_kernel<__nv_bfloat16>, grid, NF_THREADS, 0, s, to_lin(x), mean, rstd, gamma, beta, to_lin(gpad), to_lin(gextra), sh, relu, s1, s2, to_lin(dx), to_lin(gtotal));
  count_launch();
  AST_CUDA_LAUNCH_CHECK();
  return 1;
}

}  // namespace ast
