// Shared-memory-staged InstanceNorm apply / backward kernels (bulk-copy pipeline).
//
// The register-batched kernels of norm_fast.cu are latency bound: ncu shows long-scoreboard stalls on the 16-byte
// global loads with only 16 resident warps (128 registers) and ~64 KB of loads in flight per SM, which sustains about
// 3 TB/s.  Here a dedicated producer warp streams whole row segments (<= 16 KB per tensor) into a shared-memory ring
// with cp.async.bulk (the 1-D TMA copy, completion on an mbarrier), so the bytes in flight (up to ~190 KB per SM) no
// longer depend on registers; eight consumer warps read the staged rows with conflict-free 16-byte shared loads, do
// the math and write their results straight to global memory with coalesced 16-byte stores.
//
// Same math, same thread <-> channel-group mapping and same folded constants as norm_fast.cu (see there):
//   y  = A*x + D,   dx = A*g' + B*x + C,   g' = (fold_reflect(gpad) + gextra) * [A*x + D > 0]
// Requirements checked on the host: NHWC with pixel stride == C (dense rows), 16-byte aligned rows, one dtype.
#include "tc_common.cuh"

namespace ast {

constexpr int NS_CONSUMER_WARPS = 8;
constexpr int NS_CONSUMERS = NS_CONSUMER_WARPS * 32;
constexpr int NS_THREADS = NS_CONSUMERS + 32;      // + one producer warp
constexpr int NS_SLAB_MAX = 16384;                 // bytes of one tensor's row segment
constexpr int NS_MAX_STAGES = 8;
constexpr int NS_SMEM_BUDGET = 100 * 1024;         // per block: two blocks per SM

struct Rows {            // dense-row tensor: element offset of pixel (n, i, j) = n*sn + i*sh + j*C
  const char* ptr;
  int sn, sh;
};

struct NsShape {
  int C, H, W, pad;
  int nblk;              // blocks per image: block b owns rows [b*R/nblk, (b+1)*R/nblk) of the R iterated rows
  int seg, nseg;         // pixels per row segment (of the iterated row), segments per row
  int stages, slab_bytes;
};

template <typename T> struct NsVec;
template <> struct NsVec<float> {
  static constexpr int N = 4;
  __device__ static __forceinline__ void unpack(const uint4& r, float* v) {
    v[0] = __uint_as_float(r.x); v[1] = __uint_as_float(r.y); v[2] = __uint_as_float(r.z); v[3] = __uint_as_float(r.w);
  }
  __device__ static __forceinline__ uint4 pack(const float* v) {
    return make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
  }
};
template <> struct NsVec<__nv_bfloat16> {
  static constexpr int N = 8;
  __device__ static __forceinline__ void unpack(const uint4& r, float* v) {
    const unsigned w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
  }
  __device__ static __forceinline__ uint4 pack(const float* v) {
    uint4 r;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    return r;
  }
};

template <int VEC>
__device__ __forceinline__ void ns_ldc(const float* __restrict__ p, float* v) {
#pragma unroll
  for (int e = 0; e < VEC; e += 4) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p + e));
    v[e] = t.x; v[e + 1] = t.y; v[e + 2] = t.z; v[e + 3] = t.w;
  }
}

__device__ __forceinline__ void bulk_load(unsigned dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <typename T>
__device__ __forceinline__ const T* rows_ptr(const Rows& t, int off) { return reinterpret_cast<const T*>(t.ptr) + off; }
template <typename T>
__device__ __forceinline__ void rows_store(const Rows& t, int off, const uint4& v) {
  *reinterpret_cast<uint4*>(reinterpret_cast<T*>(const_cast<char*>(t.ptr)) + off) = v;
}

struct NsPipe {
  unsigned long long* full;
  unsigned long long* empty;
  unsigned buf;            // shared-space address of the ring
};

// Common prologue: barriers, then the block splits into the producer warp and the consumers.
__device__ __forceinline__ void ns_init(unsigned long long* full, unsigned long long* empty, int stages) {
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NS_CONSUMER_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
}

// x columns an output segment [oa, ob) of a reflect-padded row needs: [xa, xb)
__device__ __forceinline__ void apply_span(int oa, int ob, int pad, int W, int& xa, int& xb) {
  const int lo = oa - pad, hi = ob - 1 - pad;
  xa = max(lo, 0); xb = min(hi + 1, W);
  if (lo < 0) xb = max(xb, min(W, 1 - lo));
  if (hi >= W) xa = min(xa, max(0, 2 * (W - 1) - hi));
}

// ---------------------------------------------------------------- forward apply
template <typename T>
__global__ void __launch_bounds__(NS_THREADS, 2)
in_apply_staged_kernel(Rows x, const float* __restrict__ mean, const float* __restrict__ rstd,
                       const float* __restrict__ gamma, const float* __restrict__ beta, Rows res, Rows out, NsShape sh,
                       int relu) {
  constexpr int VEC = NsVec<T>::N;
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long full[NS_MAX_STAGES], empty[NS_MAX_STAGES];
  const unsigned buf = (smem_u32(smem_raw) + 127u) & ~127u;
  ns_init(full, empty, sh.stages);
  const int n = blockIdx.y, C = sh.C;
  const int OH = sh.H + 2 * sh.pad, OW = sh.W + 2 * sh.pad;
  const int rbeg = (int)((long long)blockIdx.x * OH / sh.nblk), rend = (int)((long long)(blockIdx.x + 1) * OH / sh.nblk);
  const int nunits = (rend - rbeg) * sh.nseg;
  const int nslabs = res.ptr ? 2 : 1;
  const int stage_bytes = nslabs * sh.slab_bytes;

  if (threadIdx.x >= NS_CONSUMERS) {
    if (threadIdx.x == NS_CONSUMERS) {
      for (int k = 0; k < nunits; ++k) {
        const int s = k % sh.stages;
        const unsigned ph = (unsigned)(k / sh.stages) & 1u;
        const int oy = rbeg + k / sh.nseg, sg = k % sh.nseg;
        const int i = reflect_idx(oy - sh.pad, sh.H);
        const int oa = sg * sh.seg, ob = min(OW, oa + sh.seg);
        int xa, xb;
        apply_span(oa, ob, sh.pad, sh.W, xa, xb);
        const unsigned bytes = (unsigned)((xb - xa) * C * (int)sizeof(T));
        mbar_wait(&empty[s], ph ^ 1u);
        mbar_expect_tx(&full[s], bytes * nslabs);
        bulk_load(buf + s * stage_bytes, rows_ptr<T>(x, n * x.sn + i * x.sh + xa * C), bytes, &full[s]);
        if (res.ptr) bulk_load(buf + s * stage_bytes + sh.slab_bytes, rows_ptr<T>(res, n * res.sn + i * res.sh + xa * C), bytes, &full[s]);
      }
    }
    return;
  }

  const int lanes = C / VEC, slots = NS_CONSUMERS / lanes;
  const int lane = threadIdx.x % lanes, slot = threadIdx.x / lanes;
  const int c = lane * VEC;
  float A[VEC], D[VEC];
  {
    float mu[VEC];
    ns_ldc<VEC>(gamma + c, A); ns_ldc<VEC>(rstd + n * C + c, D); ns_ldc<VEC>(mean + n * C + c, mu);
#pragma unroll
    for (int e = 0; e < VEC; ++e) A[e] *= D[e];
    ns_ldc<VEC>(beta + c, D);
#pragma unroll
    for (int e = 0; e < VEC; ++e) D[e] -= A[e] * mu[e];
  }
  const int ob0 = n * out.sn + c;
  for (int k = 0; k < nunits; ++k) {
    const int s = k % sh.stages;
    const unsigned ph = (unsigned)(k / sh.stages) & 1u;
    const int oy = rbeg + k / sh.nseg, sg = k % sh.nseg;
    const int oa = sg * sh.seg, ob = min(OW, oa + sh.seg);
    int xa, xb;
    apply_span(oa, ob, sh.pad, sh.W, xa, xb);
    const unsigned xs = buf + s * stage_bytes + (unsigned)(c * (int)sizeof(T));
    const int orow = ob0 + oy * out.sh;
    mbar_wait(&full[s], ph);
    for (int ox = oa + slot; ox < ob; ox += slots) {
      const int j = reflect_idx(ox - sh.pad, sh.W) - xa;
      float v[VEC];
      NsVec<T>::unpack(lds128(xs + (unsigned)(j * C * (int)sizeof(T))), v);
#pragma unroll
      for (int e = 0; e < VEC; ++e) v[e] = fmaf(A[e], v[e], D[e]);
      if (res.ptr) {
        float r[VEC];
        NsVec<T>::unpack(lds128(xs + sh.slab_bytes + (unsigned)(j * C * (int)sizeof(T))), r);
#pragma unroll
        for (int e = 0; e < VEC; ++e) v[e] += r[e];
      }
      if (relu) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) v[e] = fmaxf(v[e], 0.f);
      }
      rows_store<T>(out, orow + ox * C, NsVec<T>::pack(v));
    }
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[s]);
  }
}

// ---------------------------------------------------------------- backward
// mirrored positions of the padded gradient that fold onto (i, j): only evaluated on the border ring (global loads)
template <typename T, int VEC>
__device__ __forceinline__ void ns_fold_border(const Rows& gpad, int gbase, int C, int pad, int H, int W, int i, int j, float* g) {
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
      NsVec<T>::unpack(__ldg(reinterpret_cast<const uint4*>(rows_ptr<T>(gpad, gbase + r * gpad.sh + cc * C))), t);
#pragma unroll
      for (int e = 0; e < VEC; ++e) g[e] += t[e];
    }
  }
}

// producer of both backward kernels: x row segment, the centre of the padded gradient row, the extra gradient
template <typename T>
__device__ __forceinline__ void ns_bwd_produce(const Rows& x, const Rows& gpad, const Rows& gextra, const NsShape& sh, int n,
                                               int rbeg, int nunits, unsigned buf, int stage_bytes,
                                               unsigned long long* full, unsigned long long* empty) {
  const int C = sh.C;
  const int nslabs = 1 + (gpad.ptr ? 1 : 0) + (gextra.ptr ? 1 : 0);
  for (int k = 0; k < nunits; ++k) {
    const int s = k % sh.stages;
    const unsigned ph = (unsigned)(k / sh.stages) & 1u;
    const int i = rbeg + k / sh.nseg, sg = k % sh.nseg;
    const int ja = sg * sh.seg, jb = min(sh.W, ja + sh.seg);
    const unsigned bytes = (unsigned)((jb - ja) * C * (int)sizeof(T));
    mbar_wait(&empty[s], ph ^ 1u);
    mbar_expect_tx(&full[s], bytes * nslabs);
    unsigned dst = buf + s * stage_bytes;
    bulk_load(dst, rows_ptr<T>(x, n * x.sn + i * x.sh + ja * C), bytes, &full[s]);
    dst += sh.slab_bytes;
    if (gpad.ptr) {
      bulk_load(dst, rows_ptr<T>(gpad, n * gpad.sn + (i + sh.pad) * gpad.sh + (ja + sh.pad) * C), bytes, &full[s]);
      dst += sh.slab_bytes;
    }
    if (gextra.ptr) bulk_load(dst, rows_ptr<T>(gextra, n * gextra.sn + i * gextra.sh + ja * C), bytes, &full[s]);
  }
}

// g' of the pixel whose staged vectors start at shared addresses xs / gs / es (0 = tensor absent)
#define NS_GPRIME()                                                                                     \
  float xv[VEC], g[VEC];                                                                                \
  NsVec<T>::unpack(lds128(xs + poff), xv);                                                              \
  if (gpad.ptr) {                                                                                       \
    NsVec<T>::unpack(lds128(gs + poff), g);                                                             \
    if (sh.pad > 0 && (brow || j <= sh.pad || j >= sh.W - 1 - sh.pad))                                  \
      ns_fold_border<T, VEC>(gpad, gb, C, sh.pad, sh.H, sh.W, i, j, g);                                 \
  } else {                                                                                              \
    _Pragma("unroll") for (int e = 0; e < VEC; ++e) g[e] = 0.f;                                         \
  }                                                                                                     \
  if (gextra.ptr) {                                                                                     \
    float t[VEC];                                                                                       \
    NsVec<T>::unpack(lds128(es + poff), t);                                                             \
    _Pragma("unroll") for (int e = 0; e < VEC; ++e) g[e] += t[e];                                       \
  }                                                                                                     \
  _Pragma("unroll") for (int e = 0; e < VEC; ++e)                                                       \
    if (relu && !(fmaf(A[e], xv[e], D[e]) > 0.f)) g[e] = 0.f;   /* same expression as the forward apply */

template <typename T>
__global__ void __launch_bounds__(NS_THREADS, 2)
in_bwd_stats_staged_kernel(Rows x, const float* __restrict__ mean, const float* __restrict__ rstd,
                           const float* __restrict__ gamma, const float* __restrict__ beta, Rows gpad, Rows gextra,
                           NsShape sh, int relu, float* __restrict__ s1o, float* __restrict__ s2o) {
  constexpr int VEC = NsVec<T>::N;
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long full[NS_MAX_STAGES], empty[NS_MAX_STAGES];
  const unsigned buf = (smem_u32(smem_raw) + 127u) & ~127u;
  ns_init(full, empty, sh.stages);
  const int n = blockIdx.y, C = sh.C;
  const int rbeg = (int)((long long)blockIdx.x * sh.H / sh.nblk), rend = (int)((long long)(blockIdx.x + 1) * sh.H / sh.nblk);
  const int nunits = (rend - rbeg) * sh.nseg;
  const int nslabs = 1 + (gpad.ptr ? 1 : 0) + (gextra.ptr ? 1 : 0);
  const int stage_bytes = nslabs * sh.slab_bytes;
  const int lanes = C / VEC, slots = NS_CONSUMERS / lanes;

  if (threadIdx.x >= NS_CONSUMERS) {
    if (threadIdx.x == NS_CONSUMERS) ns_bwd_produce<T>(x, gpad, gextra, sh, n, rbeg, nunits, buf, stage_bytes, full, empty);
  } else {
    const int lane = threadIdx.x % lanes, slot = threadIdx.x / lanes;
    const int c = lane * VEC;
    float A[VEC], D[VEC], t1[VEC], t2[VEC];
    {
      float mu[VEC];
      ns_ldc<VEC>(gamma + c, A); ns_ldc<VEC>(rstd + n * C + c, D); ns_ldc<VEC>(mean + n * C + c, mu);
#pragma unroll
      for (int e = 0; e < VEC; ++e) A[e] *= D[e];
      ns_ldc<VEC>(beta + c, D);
#pragma unroll
      for (int e = 0; e < VEC; ++e) { D[e] -= A[e] * mu[e]; t1[e] = 0.f; t2[e] = 0.f; }
    }
    const int gb = n * gpad.sn + c;
    for (int k = 0; k < nunits; ++k) {
      const int s = k % sh.stages;
      const unsigned ph = (unsigned)(k / sh.stages) & 1u;
      const int i = rbeg + k / sh.nseg, sg = k % sh.nseg;
      const int ja = sg * sh.seg, jb = min(sh.W, ja + sh.seg);
      const bool brow = sh.pad > 0 && (i <= sh.pad || i >= sh.H - 1 - sh.pad);
      const unsigned xs = buf + s * stage_bytes + (unsigned)(c * (int)sizeof(T));
      const unsigned gs = xs + sh.slab_bytes;
      const unsigned es = xs + (gpad.ptr ? 2 : 1) * sh.slab_bytes;
      mbar_wait(&full[s], ph);
      for (int j = ja + slot; j < jb; j += slots) {
        const unsigned poff = (unsigned)((j - ja) * C * (int)sizeof(T));
        NS_GPRIME()
#pragma unroll
        for (int e = 0; e < VEC; ++e) { t1[e] += g[e]; t2[e] = fmaf(g[e], xv[e], t2[e]); }
      }
      __syncwarp();
      if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[s]);
    }
    // every stage has been consumed by this warp; the ring is reused as reduction scratch after the block barrier
    __syncwarp();
    float* r1 = reinterpret_cast<float*>(smem_raw + (buf - smem_u32(smem_raw)));
    float* r2 = r1 + slots * C;
    asm volatile("bar.sync 1, %0;" ::"n"(NS_CONSUMERS) : "memory");     // consumers only: all units consumed
#pragma unroll
    for (int e = 0; e < VEC; ++e) { r1[slot * C + c + e] = t1[e]; r2[slot * C + c + e] = t2[e]; }
    asm volatile("bar.sync 1, %0;" ::"n"(NS_CONSUMERS) : "memory");
    for (int cc = threadIdx.x; cc < C; cc += NS_CONSUMERS) {
      float u1 = 0.f, u2 = 0.f;
      for (int q = 0; q < slots; ++q) { u1 += r1[q * C + cc]; u2 += r2[q * C + cc]; }
      atomicAdd(s1o + n * C + cc, u1);
      atomicAdd(s2o + n * C + cc, rstd[n * C + cc] * (u2 - mean[n * C + cc] * u1));
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(NS_THREADS, 2)
in_bwd_apply_staged_kernel(Rows x, const float* __restrict__ mean, const float* __restrict__ rstd,
                           const float* __restrict__ gamma, const float* __restrict__ beta, Rows gpad, Rows gextra,
                           NsShape sh, int relu, const float* __restrict__ s1, const float* __restrict__ s2, Rows dx,
                           Rows gtotal) {
  constexpr int VEC = NsVec<T>::N;
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long full[NS_MAX_STAGES], empty[NS_MAX_STAGES];
  const unsigned buf = (smem_u32(smem_raw) + 127u) & ~127u;
  ns_init(full, empty, sh.stages);
  const int n = blockIdx.y, C = sh.C;
  const int rbeg = (int)((long long)blockIdx.x * sh.H / sh.nblk), rend = (int)((long long)(blockIdx.x + 1) * sh.H / sh.nblk);
  const int nunits = (rend - rbeg) * sh.nseg;
  const int nslabs = 1 + (gpad.ptr ? 1 : 0) + (gextra.ptr ? 1 : 0);
  const int stage_bytes = nslabs * sh.slab_bytes;

  if (threadIdx.x >= NS_CONSUMERS) {
    if (threadIdx.x == NS_CONSUMERS) ns_bwd_produce<T>(x, gpad, gextra, sh, n, rbeg, nunits, buf, stage_bytes, full, empty);
    return;
  }
  const int lanes = C / VEC, slots = NS_CONSUMERS / lanes;
  const int lane = threadIdx.x % lanes, slot = threadIdx.x / lanes;
  const int c = lane * VEC;
  const float inv_hw = 1.f / (float)(sh.H * sh.W);
  float A[VEC], B[VEC], Cc[VEC], D[VEC];
  {
    float mu[VEC], rs[VEC];
    ns_ldc<VEC>(gamma + c, A); ns_ldc<VEC>(rstd + n * C + c, rs); ns_ldc<VEC>(mean + n * C + c, mu);
    ns_ldc<VEC>(beta + c, D); ns_ldc<VEC>(s2 + n * C + c, B); ns_ldc<VEC>(s1 + n * C + c, Cc);
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      A[e] *= rs[e];
      D[e] -= A[e] * mu[e];
      B[e] = -A[e] * rs[e] * B[e] * inv_hw;
      Cc[e] = -A[e] * Cc[e] * inv_hw - B[e] * mu[e];
    }
  }
  const int gb = n * gpad.sn + c, db = n * dx.sn + c, tb = n * gtotal.sn + c;
  for (int k = 0; k < nunits; ++k) {
    const int s = k % sh.stages;
    const unsigned ph = (unsigned)(k / sh.stages) & 1u;
    const int i = rbeg + k / sh.nseg, sg = k % sh.nseg;
    const int ja = sg * sh.seg, jb = min(sh.W, ja + sh.seg);
    const bool brow = sh.pad > 0 && (i <= sh.pad || i >= sh.H - 1 - sh.pad);
    const unsigned xs = buf + s * stage_bytes + (unsigned)(c * (int)sizeof(T));
    const unsigned gs = xs + sh.slab_bytes;
    const unsigned es = xs + (gpad.ptr ? 2 : 1) * sh.slab_bytes;
    const int drow = db + i * dx.sh, trow = tb + i * gtotal.sh;
    mbar_wait(&full[s], ph);
    for (int j = ja + slot; j < jb; j += slots) {
      const unsigned poff = (unsigned)((j - ja) * C * (int)sizeof(T));
      NS_GPRIME()
      if (gtotal.ptr) rows_store<T>(gtotal, trow + j * C, NsVec<T>::pack(g));
#pragma unroll
      for (int e = 0; e < VEC; ++e) xv[e] = fmaf(A[e], g[e], fmaf(B[e], xv[e], Cc[e]));
      rows_store<T>(dx, drow + j * C, NsVec<T>::pack(xv));
    }
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[s]);
  }
}

// ---------------------------------------------------------------- backward, both passes in ONE kernel
// The two-kernel backward reads x and g' twice from HBM (7 physical passes for 5 algorithmic ones).  For layers whose
// x + g' fit in the 126 MB L2 (the thirteen 64^2 x 128 layers of the B=32 step: 67-100 MB) this kernel runs the
// statistics pass, meets the other blocks of the SAME image at a counter barrier (all blocks are co-resident: the grid is
// one wave, launched cooperatively), and re-streams its own rows for the apply pass - which now hit L2.  The producer warp
// does not wait at the barrier: it keeps prefetching the first apply units into the ring while the sums settle.
template <typename T>
__global__ void __launch_bounds__(NS_THREADS, 2)
in_bwd_fused_staged_kernel(Rows x, const float* __restrict__ mean, const float* __restrict__ rstd,
                           const float* __restrict__ gamma, const float* __restrict__ beta, Rows gpad, Rows gextra,
                           NsShape sh, int relu, float* __restrict__ s1o, float* __restrict__ s2o, int* __restrict__ arrive,
                           Rows dx, Rows gtotal) {
  constexpr int VEC = NsVec<T>::N;
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long full[NS_MAX_STAGES], empty[NS_MAX_STAGES];
  const unsigned buf = (smem_u32(smem_raw) + 127u) & ~127u;
  ns_init(full, empty, sh.stages);
  const int n = blockIdx.y, C = sh.C;
  const int rbeg = (int)((long long)blockIdx.x * sh.H / sh.nblk), rend = (int)((long long)(blockIdx.x + 1) * sh.H / sh.nblk);
  const int nunits = (rend - rbeg) * sh.nseg;
  const int nslabs = 1 + (gpad.ptr ? 1 : 0) + (gextra.ptr ? 1 : 0);
  const int stage_bytes = nslabs * sh.slab_bytes;
  const int lanes = C / VEC, slots = NS_CONSUMERS / lanes;

  if (threadIdx.x >= NS_CONSUMERS) {
    if (threadIdx.x == NS_CONSUMERS) {       // the same unit sequence twice (ring positions simply continue)
      for (int k2 = 0; k2 < 2 * nunits; ++k2) {
        const int k = k2 < nunits ? k2 : k2 - nunits;
        const int s = k2 % sh.stages;
        const unsigned ph = (unsigned)(k2 / sh.stages) & 1u;
        const int i = rbeg + k / sh.nseg, sg = k % sh.nseg;
        const int ja = sg * sh.seg, jb = min(sh.W, ja + sh.seg);
        const unsigned bytes = (unsigned)((jb - ja) * C * (int)sizeof(T));
        mbar_wait(&empty[s], ph ^ 1u);
        mbar_expect_tx(&full[s], bytes * nslabs);
        unsigned dst = buf + s * stage_bytes;
        bulk_load(dst, rows_ptr<T>(x, n * x.sn + i * x.sh + ja * C), bytes, &full[s]);
        dst += sh.slab_bytes;
        if (gpad.ptr) {
          bulk_load(dst, rows_ptr<T>(gpad, n * gpad.sn + (i + sh.pad) * gpad.sh + (ja + sh.pad) * C), bytes, &full[s]);
          dst += sh.slab_bytes;
        }
        if (gextra.ptr) bulk_load(dst, rows_ptr<T>(gextra, n * gextra.sn + i * gextra.sh + ja * C), bytes, &full[s]);
      }
    }
    return;
  }
  const int lane = threadIdx.x % lanes, slot = threadIdx.x / lanes;
  const int c = lane * VEC;
  float A[VEC], D[VEC];
  {
    float mu[VEC];
    ns_ldc<VEC>(gamma + c, A); ns_ldc<VEC>(rstd + n * C + c, D); ns_ldc<VEC>(mean + n * C + c, mu);
#pragma unroll
    for (int e = 0; e < VEC; ++e) A[e] *= D[e];
    ns_ldc<VEC>(beta + c, D);
#pragma unroll
    for (int e = 0; e < VEC; ++e) D[e] -= A[e] * mu[e];
  }
  const int gb = n * gpad.sn + c;
  // ---- pass 1: statistics
  {
    float t1[VEC], t2[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) { t1[e] = 0.f; t2[e] = 0.f; }
    for (int k = 0; k < nunits; ++k) {
      const int s = k % sh.stages;
      const unsigned ph = (unsigned)(k / sh.stages) & 1u;
      const int i = rbeg + k / sh.nseg, sg = k % sh.nseg;
      const int ja = sg * sh.seg, jb = min(sh.W, ja + sh.seg);
      const bool brow = sh.pad > 0 && (i <= sh.pad || i >= sh.H - 1 - sh.pad);
      const unsigned xs = buf + s * stage_bytes + (unsigned)(c * (int)sizeof(T));
      const unsigned gs = xs + sh.slab_bytes;
      const unsigned es = xs + (gpad.ptr ? 2 : 1) * sh.slab_bytes;
      mbar_wait(&full[s], ph);
      for (int j = ja + slot; j < jb; j += slots) {
        const unsigned poff = (unsigned)((j - ja) * C * (int)sizeof(T));
        NS_GPRIME()
#pragma unroll
        for (int e = 0; e < VEC; ++e) { t1[e] += g[e]; t2[e] = fmaf(g[e], xv[e], t2[e]); }
      }
      __syncwarp();
      if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[s]);
    }
    // block reduction in a scratch area BEHIND the ring (the producer is already refilling the ring for pass 2)
    float* r1 = reinterpret_cast<float*>(smem_raw + (buf - smem_u32(smem_raw)) + (size_t)sh.stages * stage_bytes);
    float* r2 = r1 + slots * C;
#pragma unroll
    for (int e = 0; e < VEC; ++e) { r1[slot * C + c + e] = t1[e]; r2[slot * C + c + e] = t2[e]; }
    asm volatile("bar.sync 1, %0;" ::"n"(NS_CONSUMERS) : "memory");
    for (int cc = threadIdx.x; cc < C; cc += NS_CONSUMERS) {
      float u1 = 0.f, u2 = 0.f;
      for (int q = 0; q < slots; ++q) { u1 += r1[q * C + cc]; u2 += r2[q * C + cc]; }
      atomicAdd(s1o + n * C + cc, u1);
      atomicAdd(s2o + n * C + cc, rstd[n * C + cc] * (u2 - mean[n * C + cc] * u1));
    }
    // ---- all blocks of image n have added their sums: release / acquire through the fence + counter
    __threadfence();
    asm volatile("bar.sync 1, %0;" ::"n"(NS_CONSUMERS) : "memory");
    if (threadIdx.x == 0) {
      atomicAdd(arrive + n, 1);
      const volatile int* cnt = arrive + n;
      for (unsigned spin = 0; *cnt < sh.nblk; ++spin)
        if (spin > (1u << 28)) __trap();            // bounded: a scheduling surprise traps instead of hanging the GPU
      __threadfence();
    }
    asm volatile("bar.sync 1, %0;" ::"n"(NS_CONSUMERS) : "memory");
  }
  // ---- pass 2: apply (x and g' of this block's rows were just read: they are served from L2)
  const float inv_hw = 1.f / (float)(sh.H * sh.W);
  float B[VEC], Cc[VEC];
  {
    float mu[VEC], rs[VEC];
    ns_ldc<VEC>(rstd + n * C + c, rs); ns_ldc<VEC>(mean + n * C + c, mu);
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const float S1 = __ldcg(s1o + n * C + c + e), S2 = __ldcg(s2o + n * C + c + e);
      B[e] = -A[e] * rs[e] * S2 * inv_hw;
      Cc[e] = -A[e] * S1 * inv_hw - B[e] * mu[e];
    }
  }
  const int db = n * dx.sn + c, tb = n * gtotal.sn + c;
  for (int k = 0; k < nunits; ++k) {
    const int k2 = nunits + k;
    const int s = k2 % sh.stages;
    const unsigned ph = (unsigned)(k2 / sh.stages) & 1u;
    const int i = rbeg + k / sh.nseg, sg = k % sh.nseg;
    const int ja = sg * sh.seg, jb = min(sh.W, ja + sh.seg);
    const bool brow = sh.pad > 0 && (i <= sh.pad || i >= sh.H - 1 - sh.pad);
    const unsigned xs = buf + s * stage_bytes + (unsigned)(c * (int)sizeof(T));
    const unsigned gs = xs + sh.slab_bytes;
    const unsigned es = xs + (gpad.ptr ? 2 : 1) * sh.slab_bytes;
    const int drow = db + i * dx.sh, trow = tb + i * gtotal.sh;
    mbar_wait(&full[s], ph);
    for (int j = ja + slot; j < jb; j += slots) {
      const unsigned poff = (unsigned)((j - ja) * C * (int)sizeof(T));
      NS_GPRIME()
      if (gtotal.ptr) rows_store<T>(gtotal, trow + j * C, NsVec<T>::pack(g));
#pragma unroll
      for (int e = 0; e < VEC; ++e) xv[e] = fmaf(A[e], g[e], fmaf(B[e], xv[e], Cc[e]));
      rows_store<T>(dx, drow + j * C, NsVec<T>::pack(xv));
    }
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[s]);
  }
}

// ---------------------------------------------------------------- host side
static bool rows_ok(const ast_image* im, int vec) {
  if (!im) return true;
  if (im->sc != 1 || im->sw != im->c || im->c % vec != 0 || im->sh % vec != 0 || im->sn % vec != 0 || ((uintptr_t)im->ptr & 15) != 0)
    return false;
  const long long span = (long long)(im->n - 1) * im->sn + (long long)(im->h - 1) * im->sh + (long long)im->w * im->c;
  return im->sn >= 0 && im->sh >= 0 && span < (1ll << 31);
}
static Rows to_rows(const ast_image* im) {
  Rows r;
  if (!im) { r.ptr = nullptr; r.sn = r.sh = 0; return r; }
  r.ptr = (const char*)im->ptr; r.sn = (int)im->sn; r.sh = (int)im->sh;
  return r;
}
// rows per block / segments / stages; false when the shape does not suit the staged kernels
static bool ns_plan(NsShape* sh, int n, int C, int H, int W, int pad, int rows_total, int width, int esz, int nslabs, int* nblk,
                    int slab_max = NS_SLAB_MAX, int budget = NS_SMEM_BUDGET) {
  if (C * esz > NS_SLAB_MAX / 8 || NS_CONSUMERS % (C / (16 / esz)) != 0) return false;
  sh->C = C; sh->H = H; sh->W = W; sh->pad = pad;
  const int seg_max = slab_max / (C * esz) - (width != W ? 2 * pad : 0);   // apply: the x span of a segment adds <= 2*pad
  if (seg_max < 4 * pad + 1) return false;
  sh->nseg = (width + seg_max - 1) / seg_max;
  sh->seg = (width + sh->nseg - 1) / sh->nseg;
  sh->nseg = (width + sh->seg - 1) / sh->seg;
  const int slab_px = sh->seg + (width != W ? 2 * pad : 0);
  sh->slab_bytes = (slab_px * C * esz + 127) / 128 * 128;
  sh->stages = budget / (nslabs * sh->slab_bytes);
  if (sh->stages > NS_MAX_STAGES) sh->stages = NS_MAX_STAGES;
  if (sh->stages < 2) return false;
  // as many blocks per image as fit one wave of 2 blocks per SM; rows are split proportionally (7/8 rows each at
  // H = 64, B = 32: 288 of the 296 block slots busy instead of 256 with a fixed 8 rows per block)
  int nb = (2 * num_sms()) / n;
  if (nb > rows_total) nb = rows_total;
  if (nb < 1) nb = 1;
  sh->nblk = nb;
  *nblk = nb;
  return true;
}

template <typename K>
static cudaError_t ns_attr(K kernel, size_t smem) {
  return set_max_smem(kernel, smem);
}

// Each returns 1 if the staged kernel was launched, 0 if the caller should use the register kernels.
int instnorm_apply_staged(const ast_image* x, const float* mean, const float* rstd, const float* gamma, const float* beta,
                          const ast_image* residual, const ast_image* out, int pad, int relu, cudaStream_t s) {
  const int vec = x->dtype == AST_F32 ? 4 : 8, esz = x->dtype == AST_F32 ? 4 : 2;
  if (x->dtype != out->dtype || !rows_ok(x, vec) || !rows_ok(out, vec)) return 0;
  if (residual && (residual->dtype != x->dtype || !rows_ok(residual, vec))) return 0;
  if (pad >= x->h || pad >= x->w) return 0;
  NsShape sh;
  int nblk;
  const int nslabs = residual ? 2 : 1;
  if (!ns_plan(&sh, x->n, x->c, x->h, x->w, pad, out->h, out->w, esz, nslabs, &nblk)) return 0;
  const size_t smem = (size_t)sh.stages * nslabs * sh.slab_bytes + 128;
  dim3 grid(nblk, x->n);
  cudaError_t e;
  if (x->dtype == AST_F32) {
    e = ns_attr(in_apply_staged_kernel<float>, smem);
    if (e == cudaSuccess) launch_k(in_apply_staged_kernel<float>, grid, NS_THREADS, smem, s, to_rows(x), mean, rstd, gamma, beta, to_rows(residual), to_rows(out), sh, relu);
  } else {
    e = ns_attr(in_apply_staged_kernel<__nv_bfloat16>, smem);
    if (e == cudaSuccess) launch_k(in_apply_staged_kernel<__nv_bfloat16>, grid, NS_THREADS, smem, s, to_rows(x), mean, rstd, gamma, beta, to_rows(residual), to_rows(out), sh, relu);
  }
  if (e != cudaSuccess) { set_error("instnorm_apply_staged: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); return (int)e; }
  count_launch();
  count_work(FAM_IN_APPLY, 0.0, img_bytes(x) + img_bytes(out) + img_bytes(residual));
  AST_CUDA_LAUNCH_CHECK();
  return 1;
}

static bool bwd_rows_ok(const ast_image* x, const ast_image* gpad, const ast_image* gextra, int vec) {
  if (!rows_ok(x, vec)) return false;
  if (gpad && (gpad->dtype != x->dtype || !rows_ok(gpad, vec))) return false;
  if (gextra && (gextra->dtype != x->dtype || !rows_ok(gextra, vec))) return false;
  return true;
}

int instnorm_bwd_stats_staged(const ast_image* x, const float* mean, const float* rstd, const float* gamma,
                              const float* beta, const ast_image* gpad, int pad, const ast_image* gextra, int relu,
                              float* s1, float* s2, cudaStream_t s) {
  const int vec = x->dtype == AST_F32 ? 4 : 8, esz = x->dtype == AST_F32 ? 4 : 2;
  if (!bwd_rows_ok(x, gpad, gextra, vec)) return 0;
  NsShape sh;
  int nblk;
  const int nslabs = 1 + (gpad ? 1 : 0) + (gextra ? 1 : 0);
  if (!ns_plan(&sh, x->n, x->c, x->h, x->w, pad, x->h, x->w, esz, nslabs, &nblk)) return 0;
  const size_t scratch = 2 * (size_t)(NS_CONSUMERS / (x->c / vec)) * x->c * sizeof(float);
  size_t smem = (size_t)sh.stages * nslabs * sh.slab_bytes;
  if (smem < scratch) smem = scratch;
  smem += 128;
  dim3 grid(nblk, x->n);
  cudaError_t e;
  if (x->dtype == AST_F32) {
    e = ns_attr(in_bwd_stats_staged_kernel<float>, smem);
    if (e == cudaSuccess) launch_k(in_bwd_stats_staged_kernel<float>, grid, NS_THREADS, smem, s, to_rows(x), mean, rstd, gamma, beta, to_rows(gpad), to_rows(gextra), sh, relu, s1, s2);
  } else {
    e = ns_attr(in_bwd_stats_staged_kernel<__nv_bfloat16>, smem);
    if (e == cudaSuccess) launch_k(in_bwd_stats_staged_kernel<__nv_bfloat16>, grid, NS_THREADS, smem, s, to_rows(x), mean, rstd, gamma, beta, to_rows(gpad), to_rows(gextra), sh, relu, s1, s2);
  }
  if (e != cudaSuccess) { set_error("instnorm_bwd_stats_staged: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); return (int)e; }
  count_launch();
  count_work(FAM_IN_BWD, 0.0, 0.0);     // algorithmic bytes of the backward pair are booked by the apply launch
  AST_CUDA_LAUNCH_CHECK();
  return 1;
}

int instnorm_bwd_apply_staged(const ast_image* x, const float* mean, const float* rstd, const float* gamma,
                              const float* beta, const ast_image* gpad, int pad, const ast_image* gextra, int relu,
                              const float* s1, const float* s2, const ast_image* dx, const ast_image* gtotal,
                              cudaStream_t s) {
  const int vec = x->dtype == AST_F32 ? 4 : 8, esz = x->dtype == AST_F32 ? 4 : 2;
  if (x->dtype != dx->dtype || !bwd_rows_ok(x, gpad, gextra, vec) || !rows_ok(dx, vec)) return 0;
  if (gtotal && (gtotal->dtype != x->dtype || !rows_ok(gtotal, vec))) return 0;
  NsShape sh;
  int nblk;
  const int nslabs = 1 + (gpad ? 1 : 0) + (gextra ? 1 : 0);
  if (!ns_plan(&sh, x->n, x->c, x->h, x->w, pad, x->h, x->w, esz, nslabs, &nblk)) return 0;
  const size_t smem = (size_t)sh.stages * nslabs * sh.slab_bytes + 128;
  dim3 grid(nblk, x->n);
  cudaError_t e;
  if (x->dtype == AST_F32) {
    e = ns_attr(in_bwd_apply_staged_kernel<float>, smem);
    if (e == cudaSuccess) launch_k(in_bwd_apply_staged_kernel<float>, grid, NS_THREADS, smem, s, to_rows(x), mean, rstd, gamma, beta, to_rows(gpad), to_rows(gextra), sh, relu, s1, s2, to_rows(dx), to_rows(gtotal));
  } else {
    e = ns_attr(in_bwd_apply_staged_kernel<__nv_bfloat16>, smem);
    if (e == cudaSuccess) launch_k(in_bwd_apply_staged_kernel<__nv_bfloat16>, grid, NS_THREADS, smem, s, to_rows(x), mean, rstd, gamma, beta, to_rows(gpad), to_rows(gextra), sh, relu, s1, s2, to_rows(dx), to_rows(gtotal));
  }
  if (e != cudaSuccess) { set_error("instnorm_bwd_apply_staged: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); return (int)e; }
  count_launch();
  // algorithmic traffic of the InstanceNorm backward: read x and g' once, write dx (+ the skip gradient)
  count_work(FAM_IN_BWD, 0.0, 2.0 * img_bytes(x) + img_bytes(dx) + img_bytes(gtotal));
  AST_CUDA_LAUNCH_CHECK();
  return 1;
}

// (Measured and not kept: walking the batch in L2-sized image groups inside this kernel so that the 256^2 x 32 and
// 128^2 x 64 layers qualify as well - 209 vs 164 us and 84 vs 84 us against the two-kernel path: the per-group barriers and
// the short per-block row ranges cost more than the second HBM read they save.)
// Both backward passes in one cooperative launch when x + g' can stay in L2 between them.  `arrive`: n ints, zero on entry.
// Returns 1 = launched, 0 = not applicable (caller runs the two-kernel path), other = error.
int instnorm_bwd_fused_staged(const ast_image* x, const float* mean, const float* rstd, const float* gamma,
                              const float* beta, const ast_image* gpad, int pad, const ast_image* gextra, int relu,
                              float* s1, float* s2, int* arrive, const ast_image* dx, const ast_image* gtotal,
                              cudaStream_t s) {
  const int vec = x->dtype == AST_F32 ? 4 : 8, esz = x->dtype == AST_F32 ? 4 : 2;
  if (x->dtype != dx->dtype || !bwd_rows_ok(x, gpad, gextra, vec) || !rows_ok(dx, vec)) return 0;
  if (gtotal && (gtotal->dtype != x->dtype || !rows_ok(gtotal, vec))) return 0;
  const int nslabs = 1 + (gpad ? 1 : 0) + (gextra ? 1 : 0);
  const double resident = (double)nslabs * x->n * x->h * x->w * x->c * esz;      // bytes re-read by the second pass
  if (resident > 104e6) return 0;                    // would not survive in the 126 MB L2 next to the dx / skip-gradient writes
  NsShape sh;
  int nblk;
  const size_t scratch = 2 * (size_t)(NS_CONSUMERS / (x->c / vec)) * x->c * sizeof(float);
  // the reduction scratch sits behind the ring (the producer refills the ring during the barrier): give the ring what is
  // left of the per-block budget, with 8 KB slabs when three tensors are staged
  if (!ns_plan(&sh, x->n, x->c, x->h, x->w, pad, x->h, x->w, esz, nslabs, &nblk, nslabs == 3 ? 8192 : NS_SLAB_MAX,
               NS_SMEM_BUDGET + 8192 - (int)scratch)) return 0;
  if ((long long)nblk * x->n > 2ll * num_sms()) return 0;          // the counter barrier needs every block resident
  const size_t smem = (size_t)sh.stages * nslabs * sh.slab_bytes + scratch + 128;
  if (smem > 110 * 1024) return 0;
  dim3 grid(nblk, x->n);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = dim3(NS_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeCooperative;
  attr.val.cooperative = 1;
  cfg.attrs = &attr; cfg.numAttrs = 1;
  cudaError_t e;
  Rows xr = to_rows(x), gp = to_rows(gpad), ge = to_rows(gextra), dxr = to_rows(dx), gt = to_rows(gtotal);
  if (x->dtype == AST_F32) {
    e = ns_attr(in_bwd_fused_staged_kernel<float>, smem);
    if (e == cudaSuccess) e = cudaLaunchKernelEx(&cfg, in_bwd_fused_staged_kernel<float>, xr, mean, rstd, gamma, beta, gp, ge, sh, relu, s1, s2, arrive, dxr, gt);
  } else {
    e = ns_attr(in_bwd_fused_staged_kernel<__nv_bfloat16>, smem);
    if (e == cudaSuccess) e = cudaLaunchKernelEx(&cfg, in_bwd_fused_staged_kernel<__nv_bfloat16>, xr, mean, rstd, gamma, beta, gp, ge, sh, relu, s1, s2, arrive, dxr, gt);
  }
  if (e != cudaSuccess) { set_error("instnorm_bwd_fused_staged: launch failed: %s", cudaGetErrorString(e)); return (int)e; }
  count_launch();
  count_work(FAM_IN_BWD, 0.0, 2.0 * img_bytes(x) + img_bytes(dx) + img_bytes(gtotal));
  AST_CUDA_LAUNCH_CHECK();
  return 1;
}

}  // namespace ast
