// tcgen05 / TMEM / TMA pixel-contraction kernel for sm_100a: both operands are NHWC activations, i.e. MN-major
// (channels contiguous) with the reduction running over pixels:
//
//   D[m][n] (fp32, TMEM, 128 x BN)  +=  sum_{pixels p in chunk}  R[p'(p)][m0+m] * C[p''(p)][n0+n]
//
//   * filter gradient of a gather convolution (aten::convolution_backward, train_cnn.py:333):
//         R = dY at (oy0+so*i, ox0+so*j),  C = X at (si*i+dy_t, si*j+dx_t),  one CTA column per tap t
//   * Gram matrix gram() (train_cnn.py:103-107):  R = C = F, one output per image, scale 1/(CHW)
// Split-K over pixel chunks across CTAs, fp32 atomics into the (pre-zeroed) output.
// Warp roles as in conv_tc.cu: warp 0 TMA producer (4-D boxes, 128B swizzle, OOB = zero padding, elementStrides =
// coordinate multiplier), warp 1 tcgen05.mma issuer (MN-major A and B descriptors), warps 2-5 epilogue.
#include "tc_common.cuh"

namespace ast {

constexpr int CT_MAX_PROD = 4;                  // TMA producer warps: warp 0 and warps 6..8
constexpr int CT_THREADS = 192 + 32 * (CT_MAX_PROD - 1);
// A thread pays ~500 cycles per TMA instruction (wait + expect_tx + issue), and the filter-gradient stages need 2 to 10
// of them for 8 to 16 MMAs: with one producer the 32/64-channel layers ran at ~500 clk per MMA.  The stages are dealt
// round-robin to `nprod` producer WARPS (lanes of one warp do not help: divergent spin loops time-slice each other);
// nprod divides the stage count, so consecutive uses of one stage belong to the same producer and the 1-bit mbarrier
// parity cannot alias.
// (Measured and removed: a "halo" mode loading ONE (th + hy) x (tw + hx) patch per 64-channel box for all taps of a CTA
// and addressing the taps through shifted MN-major descriptors.  It halves the L2 -> shared memory traffic of the 3x3
// filter gradients but ran 1.8x SLOWER (residual layers 61 -> 107 us, results identical).  A second variant with ALIGNED
// shifts only (4 x 16 pixel tiles, the three taps of one filter column sharing a (4 + 2) x 16 patch, every operand start
// on a 2 KB boundary) was slower as well (61 -> 72 us): these kernels are bound by the MN-major MMAs themselves (both
// variants issue one N = 128 instruction per tap and K step instead of N = 256 + N = 128), not by the L2 -> shared
// memory traffic.)
__device__ __forceinline__ int ct_producer_index(int warp) { return warp == 0 ? 0 : (warp >= 6 ? warp - 5 : -1); }
// most producers first, then most stages: gives up at most two stages to make the stage count divisible
inline void ct_pick_producers(int& stages, int& nprod) {
  int best_d = 1, best_s = stages;
  for (int st = stages; st >= 2 && st >= stages - 2; --st)
    for (int d = CT_MAX_PROD; d >= 2; --d)
      if (st % d == 0 && d > best_d) { best_d = d; best_s = st; }
  stages = best_s; nprod = best_d;
}
constexpr int CT_MAX_STAGES = 8;

struct CtParams {
  int mi, mj, tw, th, tiles_i, tiles_j, n_img, kp;
  int r_s, r_oy, r_ox, c_s;
  int ntaps, cb, m_boxes, n_boxes, bn;
  int m_valid, n_valid, m_blocks, n_blocks;
  int ksplit, chunks_per_cta, per_img;
  long long chunks_total;           // per image when per_img, else over all images
  long long out_img_stride, s_m, s_n;
  float scale;
  int stages, nprod, box_bytes, stage_bytes, umma_k_bytes, kmma, upper_only;
  int tg, ngroups;                  // taps handled by one CTA (rows operand loaded once per stage for all of them)
  int same;                         // Gram: rows and cols are the SAME tensor at the same pixels (see shared_r below)
  unsigned sbo, layout_type;
  unsigned idesc;
  short dy[AST_MAX_TAPS];
  short dx[AST_MAX_TAPS];
};

// MN-major, 128B-swizzled operand: 64-element (128 B) column blocks LBO apart, 8-row K groups SBO (=1024 B) apart
// 16-bit types: SWIZZLE_128B (layout 2), swizzle atom = 8 K rows (SBO 1024 B).
// 32-bit types (tf32): MN-major operands need the 32-byte-atom flavour, SWIZZLE_128B_BASE32B (layout 1), whose atom
// is 4 K rows (SBO 512 B); the TMA map uses CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B to match.
__device__ __forceinline__ unsigned long long make_mn_desc(unsigned saddr, unsigned lbo_bytes, unsigned sbo_bytes,
                                                           unsigned layout_type) {
  unsigned long long d = 0;
  d |= (unsigned long long)((saddr & 0x3FFFFu) >> 4);
  d |= (unsigned long long)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (unsigned long long)(sbo_bytes >> 4) << 32;
  d |= (unsigned long long)1 << 46;
  d |= (unsigned long long)layout_type << 61;
  return d;
}

// Gram finishing step (north_star: "fuse the 1/(CHW) scale and the style-MSE reduction into the epilogue"): the LAST of the
// split-K CTAs of one output block (atomic ticket) mirrors the block into the lower triangle, accumulates
//   loss += loss_scale * sum_{i,j} (G_ij - S_ij)^2        (upper triangle, off-diagonal entries counted twice)
// and writes D = d_scale * (G - S), the symmetric per-image 1x1 weights of the Gram backward (TF32-rounded).
struct GramFin {
  int enabled;
  int* counters;                 // [n_img][m_blocks * n_blocks], zero on entry
  const float* target;           // S, or null (then only the mirror is done)
  long long target_img_stride;   // 0 = one target shared by the batch (train_cnn.py:187-190 expands one style Gram)
  double* loss; float loss_scale;   // fp64 accumulator: hundreds of small fp32 addends into a ~1e4 total would be truncated
  float* dmat; float d_scale;    // null = no gradient requested
};

template <int KIND>
__global__ void __launch_bounds__(CT_THREADS, 1)
contract_tc_kernel(const __grid_constant__ CUtensorMap tm_r, const __grid_constant__ CUtensorMap tm_c, const CtParams p,
                   float* __restrict__ out, const int* __restrict__ tap_off, const GramFin fin) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long full_bar[CT_MAX_STAGES], empty_bar[CT_MAX_STAGES], tfull_bar;
  __shared__ unsigned tmem_slot;
  __shared__ int s_last;
  __shared__ float s_red[4];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t0 = blockIdx.y * p.tg;                                   // first tap of this CTA's tap group
  const int nt = min(p.tg, p.ntaps - t0);
  const int mb_idx = blockIdx.z / p.n_blocks, nb_idx = blockIdx.z % p.n_blocks;
  if (p.upper_only && nb_idx * p.bn + p.bn <= mb_idx * 128) return;   // block strictly below the diagonal
  int img_fixed = -1;
  long long cbeg, cend;
  if (p.per_img) {
    img_fixed = blockIdx.x / p.ksplit;
    const int kx = blockIdx.x % p.ksplit;
    cbeg = (long long)kx * p.chunks_per_cta;
  } else {
    cbeg = (long long)blockIdx.x * p.chunks_per_cta;
  }
  cend = cbeg + p.chunks_per_cta;
  if (cend > p.chunks_total) cend = p.chunks_total;
  if (cbeg >= cend) return;                          // uniform across the CTA: nothing to do

  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int acc_cols = p.tg * p.bn;
  const unsigned tmem_cols = acc_cols <= 32 ? 32 : acc_cols <= 64 ? 64 : acc_cols <= 128 ? 128 : acc_cols <= 256 ? 256 : 512;
  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_r) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_c) : "memory");
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&tfull_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem_base = tmem_slot;
  const int m0 = mb_idx * 128, n0 = nb_idx * p.bn;
  const int r_bytes = p.m_boxes * p.box_bytes;
  // Gram blocks whose (valid) row channels lie inside the column channel range read the A operand straight out of the
  // column boxes: no second TMA load of the same pixels (rows past m_valid read whatever follows and are ignored)
  const bool shared_r = p.same && m0 >= n0 && min(m0 + 128, p.m_valid) <= n0 + p.bn;

  const int prod = ct_producer_index(warp);
  if (prod >= 0) {
    if (prod < p.nprod) {
    const int per_img_chunks = p.tiles_i * p.tiles_j;
    for (long long c = cbeg + prod; c < cend; c += p.nprod) {
      const int s = (int)((c - cbeg) % p.stages);
      const unsigned ph = (unsigned)(((c - cbeg) / p.stages) & 1);
      int img, rem;
      if (p.per_img) { img = img_fixed; rem = (int)c; }
      else { img = (int)(c / per_img_chunks); rem = (int)(c % per_img_chunks); }
      const int ti = rem / p.tiles_j, tj = rem % p.tiles_j;
      const int i0 = ti * p.th, j0 = tj * p.tw;
      mbar_wait(&empty_bar[s], ph ^ 1);
      if (lane == 0) {
        unsigned char* sr = smem + (size_t)s * p.stage_bytes;
        mbar_expect_tx(&full_bar[s], (unsigned)(((shared_r ? 0 : p.m_boxes) + nt * p.n_boxes) * p.box_bytes));
        if (!shared_r)
          for (int b = 0; b < p.m_boxes; ++b)
            tma_load_4d(sr + b * p.box_bytes, &tm_r, &full_bar[s], m0 + b * p.cb, p.r_s * j0 + p.r_ox, p.r_s * i0 + p.r_oy, img);
        for (int u = 0; u < nt; ++u)
          for (int b = 0; b < p.n_boxes; ++b)
            tma_load_4d(sr + r_bytes + (u * p.n_boxes + b) * p.box_bytes, &tm_c, &full_bar[s], n0 + b * p.cb,
                        p.c_s * j0 + p.dx[t0 + u], p.c_s * i0 + p.dy[t0 + u], img);
      }
      __syncwarp();
    }
    }
  } else if (warp == 1) {
    int s = 0; unsigned ph = 0;
    bool first = true;
    for (long long c = cbeg; c < cend; ++c) {
      mbar_wait(&full_bar[s], ph);
      tc_fence_after();
      if (lane == 0) {
        const unsigned stage_addr = smem_u32(smem + (size_t)s * p.stage_bytes);
        const unsigned a_addr = shared_r ? stage_addr + r_bytes + (unsigned)((m0 - n0) / p.cb) * p.box_bytes : stage_addr;
        // consecutive taps sit in consecutive column blocks (LBO apart) and in consecutive TMEM columns, so up to
        // 256 / bn taps go into ONE MMA: an N = 256 instruction costs 128 cycles, an N <= 128 one ~104 (mma_bench)
        const int gsz = p.bn <= 128 ? 256 / p.bn : 1;
        for (int u = 0; u < nt; u += gsz) {
          const int nu = min(gsz, nt - u);
          const unsigned idesc = (p.idesc & ~(0x3Fu << 17)) | ((unsigned)((nu * p.bn) >> 3) << 17);
          const unsigned b_addr = stage_addr + r_bytes + u * p.n_boxes * p.box_bytes;
          for (int k = 0; k < p.kmma; ++k) {
            const unsigned long long ad = make_mn_desc(a_addr + k * p.umma_k_bytes, p.box_bytes, p.sbo, p.layout_type);
            const unsigned long long bd = make_mn_desc(b_addr + k * p.umma_k_bytes, p.box_bytes, p.sbo, p.layout_type);
            tc_mma<KIND>(tmem_base + (unsigned)(u * p.bn), ad, bd, idesc, (first && k == 0) ? 0u : 1u);
          }
        }
        tc_commit(&empty_bar[s]);
        if (c == cend - 1) tc_commit(&tfull_bar);
      }
      first = false;
      __syncwarp();
      if (++s == p.stages) { s = 0; ph ^= 1; }
    }
  } else {
    const int q = warp & 3;
    const int m = m0 + q * 32 + lane;
    mbar_wait(&tfull_bar, 0);
    tc_fence_after();
    // the split-K CTAs of one output all finish together: rotate each CTA's walk over (tap, column block) so that they
    // do not all hit the same addresses at the same moment
    const int nblk = p.bn / 32;
    if (fin.enabled && p.ksplit == 1) {
      // ---- single writer (no split-K): the Gram block is final in TMEM, so the whole style term is computed from the
      // registers of the tcgen05.ld - scale, G stores, (G - S)^2, D - without atomics, tickets or a second pass.  A thread
      // owns row m and 32 consecutive columns: the (m, n) stores are 128-byte runs per thread; for the mirrored (n, m)
      // entries the 32 lanes of a warp hold 32 consecutive m of one n, i.e. one coalesced 128-byte store per column.
      const int C = p.n_valid;
      float* G = out + (long long)img_fixed * p.out_img_stride;
      const float* S = fin.target ? fin.target + (long long)img_fixed * fin.target_img_stride : nullptr;
      float* D = fin.dmat ? fin.dmat + (long long)img_fixed * p.out_img_stride : nullptr;
      float part = 0.f;
      for (int cb0 = 0; cb0 < nblk; ++cb0) {
        const int nb0 = n0 + cb0 * 32;
        if (nb0 + 31 < m0 + q * 32 || nb0 >= C) continue;          // warp-uniform: chunk strictly below this warp's rows
        float v[32], sv[32];
        tc_ld32(tmem_base + ((unsigned)(q * 32) << 16) + (unsigned)(cb0 * 32), v);
        const bool row_ok = m < p.m_valid;
        if (S) {
#pragma unroll
          for (int e = 0; e < 32; e += 4) {
            float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row_ok && nb0 + e + 3 < C) t4 = __ldg(reinterpret_cast<const float4*>(S + (long long)m * C + nb0 + e));
            sv[e] = t4.x; sv[e + 1] = t4.y; sv[e + 2] = t4.z; sv[e + 3] = t4.w;
          }
        }
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int n = nb0 + e;
          const float gv = v[e] * p.scale;
          const bool up = row_ok && n < C && n >= m;
          float dv = 0.f;
          if (S && up) { dv = gv - sv[e]; part = fmaf(n > m ? 2.f : 1.f, dv * dv, part); }
          v[e] = gv;
          sv[e] = round_tf32(fin.d_scale * dv);
          if (up && n > m) {                                       // mirrored entry: lanes = consecutive m
            G[(long long)n * C + m] = gv;
            if (D) D[(long long)n * C + m] = sv[e];
          }
        }
        if (row_ok) {
#pragma unroll
          for (int e = 0; e < 32; e += 4) {
            const int n = nb0 + e;
            if (n + 3 < C && n >= m) {                             // whole quad on / above the diagonal
              *reinterpret_cast<float4*>(G + (long long)m * C + n) = make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]);
              if (D) *reinterpret_cast<float4*>(D + (long long)m * C + n) = make_float4(sv[e], sv[e + 1], sv[e + 2], sv[e + 3]);
            } else {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                if (n + k < C && n + k >= m) { G[(long long)m * C + n + k] = v[e + k]; if (D) D[(long long)m * C + n + k] = sv[e + k]; }
            }
          }
        }
      }
      if (S && fin.loss) {
        part = warp_sum(part);
        if (lane == 0) s_red[warp - 2] = part;
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (threadIdx.x == 64) atomicAdd(fin.loss, (double)fin.loss_scale * (double)(s_red[0] + s_red[1] + s_red[2] + s_red[3]));
      }
    } else {
    const int rot = (int)(blockIdx.x % (unsigned)(nt * nblk));
    for (int it = 0; it < nt * nblk; ++it) {
      const int idx = (it + rot) % (nt * nblk);
      const int u = idx / nblk;
      float* o = out + (p.per_img ? (long long)img_fixed * p.out_img_stride : 0) + (tap_off ? tap_off[t0 + u] : 0);
      const bool vec_ok = p.s_n == 1 && (p.s_m & 3) == 0 && (((uintptr_t)o) & 15) == 0;
      {
        const int c0 = (idx - u * nblk) * 32;
        float v[32];
        tc_ld32(tmem_base + ((unsigned)(q * 32) << 16) + (unsigned)(u * p.bn + c0), v);
        if (m < p.m_valid) {
          float* orow = o + (long long)m * p.s_m;
          if (vec_ok) {          // 16-byte vector reductions: 4x fewer L2 atomic operations than scalar atomicAdd
#pragma unroll
            for (int e = 0; e < 32; e += 4) {
              const int n = n0 + c0 + e;
              if (n + 3 < p.n_valid && !(p.upper_only && n < m)) {
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(orow + n), "f"(v[e] * p.scale),
                             "f"(v[e + 1] * p.scale), "f"(v[e + 2] * p.scale), "f"(v[e + 3] * p.scale) : "memory");
              } else {
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4)
                  if (n + q4 < p.n_valid && !(p.upper_only && n + q4 < m)) atomicAdd(orow + n + q4, v[e + q4] * p.scale);
              }
            }
          } else {
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              const int n = n0 + c0 + e;
              if (n < p.n_valid && !(p.upper_only && n < m)) atomicAdd(orow + (long long)n * p.s_n, v[e] * p.scale);
            }
          }
        }
      }
    }
    if (fin.enabled) {
      // ---- ticket: the last split-K CTA of this (image, block) sees every contribution (release/acquire through the
      // fence + atomic; the block is then read with ld.global.cg, i.e. from L2 where the reductions landed)
      __threadfence();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (threadIdx.x == 64) {
        const int prev = atomicAdd(fin.counters + (long long)img_fixed * AST_GRAM_COUNTERS_PER_IMAGE + blockIdx.z, 1);
        s_last = prev == p.ksplit - 1;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (s_last) {
        __threadfence();
        const int C = p.n_valid;
        float* G = out + (long long)img_fixed * p.out_img_stride;
        const float* S = fin.target ? fin.target + (long long)img_fixed * fin.target_img_stride : nullptr;
        float* D = fin.dmat ? fin.dmat + (long long)img_fixed * p.out_img_stride : nullptr;
        // 32 x 32 tiles of the block, one warp each; transposed copies go through a padded shared-memory tile (the
        // pipeline stages are idle by now) so that both the (m, n) and the (n, m) stores are coalesced
        float* tg = reinterpret_cast<float*>(smem) + (warp - 2) * (2 * 32 * 33);
        float* td = tg + 32 * 33;
        float part = 0.f;
        const int tiles_n = p.bn / 32;
        for (int tile = warp - 2; tile < 4 * tiles_n; tile += 4) {
          const int tr = tile / tiles_n, tc = tile - tr * tiles_n;
          const int mb0 = m0 + tr * 32, nb0 = n0 + tc * 32;
          if (mb0 >= p.m_valid || nb0 >= C || nb0 + 31 < mb0) continue;     // tile outside the matrix / strictly below the diagonal
          const int n = nb0 + lane;
          // all 32 + 32 loads of the tile first (independent, in flight together), then the arithmetic and the stores:
          // interleaving them serialises on the load latency because the stores may alias the loads
          float gv[32], sv[32];
#pragma unroll
          for (int r = 0; r < 32; ++r) {
            const int mm = mb0 + r;
            const bool ok = mm < p.m_valid && n < C && n >= mm;
            gv[r] = ok ? __ldcg(G + (long long)mm * C + n) : 0.f;
            sv[r] = (ok && S) ? __ldg(S + (long long)mm * C + n) : 0.f;
          }
#pragma unroll
          for (int r = 0; r < 32; ++r) {
            const int mm = mb0 + r;
            float dv = 0.f;
            if (S && mm < p.m_valid && n < C && n >= mm) {
              dv = gv[r] - sv[r];
              part = fmaf(n > mm ? 2.f : 1.f, dv * dv, part);
              if (D) D[(long long)mm * C + n] = round_tf32(fin.d_scale * dv);
            }
            tg[r * 33 + lane] = gv[r];
            td[r * 33 + lane] = dv;
          }
          __syncwarp();
          // transposed: row n' = nb0 + r, column m' = mb0 + lane  (strictly lower entries only)
#pragma unroll 4
          for (int r = 0; r < 32; ++r) {
            const int nn = nb0 + r, mm = mb0 + lane;
            if (nn < C && mm < p.m_valid && nn > mm) {
              G[(long long)nn * C + mm] = tg[lane * 33 + r];
              if (D) D[(long long)nn * C + mm] = round_tf32(fin.d_scale * td[lane * 33 + r]);
            }
          }
          __syncwarp();
        }
        if (S && fin.loss) {
          part = warp_sum(part);
          if (lane == 0) s_red[warp - 2] = part;
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (threadIdx.x == 64) atomicAdd(fin.loss, (double)fin.loss_scale * (double)(s_red[0] + s_red[1] + s_red[2] + s_red[3]));
        }
      }
    }
    }   // split-K (ticket) path
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// fills the strict lower triangle of each C x C matrix from the upper one
__global__ void mirror_upper_kernel(float* __restrict__ g, int c, long long total) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(idx % c);
    const int i = (int)((idx / c) % c);
    if (j < i) g[idx] = g[idx - (long long)i * c - j + (long long)j * c + i];
  }
}

static int encode_operand(EncodeTiledFn encode, CUtensorMap* tm, const ast_image* im, int cb, int tw, int th, int s) {
  const int esz = im->dtype == AST_F32 ? 4 : 2;
  const CUtensorMapDataType dt = im->dtype == AST_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  cuuint64_t dims[4] = {(cuuint64_t)im->c, (cuuint64_t)im->w, (cuuint64_t)im->h, (cuuint64_t)im->n};
  cuuint64_t strides[3] = {(cuuint64_t)im->sw * esz, (cuuint64_t)im->sh * esz, (cuuint64_t)im->sn * esz};
  cuuint32_t box[4] = {(cuuint32_t)cb, (cuuint32_t)(tw * s), (cuuint32_t)(th * s), 1};
  cuuint32_t estr[4] = {1, (cuuint32_t)s, (cuuint32_t)s, 1};
  return cached_tensor_map(encode, tm, dt, 4, im->ptr, dims, strides, box, estr,
                           im->dtype == AST_F32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
}

static int check_operand(const char* who, const ast_image* im) {
  const int esz = im->dtype == AST_F32 ? 4 : 2;
  AST_CHECK_ARG(im->sc == 1, "%s: NHWC operand required", who);
  AST_CHECK_ARG(((uintptr_t)im->ptr & 15) == 0 && (im->sw * esz) % 16 == 0 && (im->sh * esz) % 16 == 0 && (im->sn * esz) % 16 == 0,
                "%s: pointer/strides must be 16-byte aligned (c=%d)", who, im->c);
  return 0;
}


// ---------------------------------------------------------------------------------------------------------------------
// Thin variant (both operands have <= 32 bf16 channels: the 9x9 first/last layers after the row unfold).  The generic
// kernel pads such operands to 64-channel boxes and to M = 128, i.e. 8x wasted MMA work and 2.6x wasted smem fill.
// Here a box is 64 pixels x 32 channels (64-byte rows, SWIZZLE_64B) and the M = 128 rows of one MMA are FOUR TAPS of the
// shifted operand (four boxes, LBO apart), the N = 32 columns the fixed operand:
//     D_g[(t - 4g)*32 + cc][rc] += sum_p C[p + tap_t][cc] * R[p][rc]            g = t / 4
// One CTA handles all taps (ceil(ntaps/4) accumulators of 32 TMEM columns); split-K over pixel chunks, fp32 atomics.
struct CtThinParams {
  int mi, mj, tw, th, tiles_i, tiles_j, n_img;
  int r_s, r_oy, r_ox, c_s;
  int ntaps, ngrp, r_valid, c_valid, chunks_per_cta, stages, nprod, stage_bytes;
  int vstack;                       // all taps are consecutive rows of one column and a chunk is one 64-pixel row segment:
                                    // ONE box of ntaps rows replaces ntaps loads (tap t = 4 KB further into it)
  int merged;                       // vstack + unit row multiplier: ONE MMA per K step for all taps (see contract_thin)
  int r_rows;                       // merged: rows of the fixed operand's patch = 4 * (ngrp - 1) + 1
  long long chunks_total, s_m, s_n;
  float scale;
  unsigned idesc;
  short dy[AST_MAX_TAPS];
  short dx[AST_MAX_TAPS];
};
constexpr int THIN_KP = 64;                    // pixels per stage
constexpr int THIN_BOX = THIN_KP * 64;         // bytes per box

__global__ void __launch_bounds__(CT_THREADS, 1)
contract_thin_kernel(const __grid_constant__ CUtensorMap tm_r, const __grid_constant__ CUtensorMap tm_c,
                     const CtThinParams p, float* __restrict__ out, const int* __restrict__ tap_off) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long full_bar[CT_MAX_STAGES], empty_bar[CT_MAX_STAGES], tfull_bar;
  __shared__ unsigned tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long cbeg = (long long)blockIdx.x * p.chunks_per_cta;
  long long cend = cbeg + p.chunks_per_cta;
  if (cend > p.chunks_total) cend = p.chunks_total;
  if (cbeg >= cend) return;

  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const unsigned tmem_cols = p.ngrp * 32 <= 32 ? 32 : p.ngrp * 32 <= 64 ? 64 : p.ngrp * 32 <= 128 ? 128 : p.ngrp * 32 <= 256 ? 256 : 512;
  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_r) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_c) : "memory");
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&tfull_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem_base = tmem_slot;
  const int padded_taps = p.ngrp * 4;          // box slots of the shifted operand per stage (slots >= ntaps stay unwritten)

  const int prod = ct_producer_index(warp);
  if (prod >= 0) {
    const int per_img_chunks = p.tiles_i * p.tiles_j;
    if (warp == 0 && lane == 0 && !p.merged) {   // unwritten tap slots feed only ignored accumulator rows, but keep them finite: zero once
      for (int st = 0; st < p.stages; ++st)
        for (int t = p.ntaps; t < padded_taps; ++t) {
          uint4* z = (uint4*)(smem + (size_t)st * p.stage_bytes + (1 + t) * THIN_BOX);
          for (int k = 0; k < THIN_BOX / 16; ++k) z[k] = make_uint4(0, 0, 0, 0);
        }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    if (prod < p.nprod) {
    for (long long c = cbeg + prod; c < cend; c += p.nprod) {
      const int s = (int)((c - cbeg) % p.stages);
      const unsigned ph = (unsigned)(((c - cbeg) / p.stages) & 1);
      const int img = (int)(c / per_img_chunks), rem = (int)(c % per_img_chunks);
      const int ti = rem / p.tiles_j, tj = rem % p.tiles_j;
      const int i0 = ti * p.th, j0 = tj * p.tw;
      mbar_wait(&empty_bar[s], ph ^ 1);
      if (lane == 0) {
        unsigned char* sr = smem + (size_t)s * p.stage_bytes;
        if (p.merged) {      // fixed operand: rows i0 - (r_rows - 1) .. i0 (rows outside the image read 0); shifted: taps 0..3
          mbar_expect_tx(&full_bar[s], (unsigned)((p.r_rows + 4 + 2 * (p.th - 1)) * THIN_BOX));
          tma_load_4d(sr, &tm_r, &full_bar[s], 0, j0 + p.r_ox, i0 - (p.r_rows - 1) + p.r_oy, img);
          tma_load_4d(sr + (p.r_rows + p.th - 1) * THIN_BOX, &tm_c, &full_bar[s], 0, j0 + p.dx[0], i0 + p.dy[0], img);
        } else {
          mbar_expect_tx(&full_bar[s], (unsigned)((1 + p.ntaps) * THIN_BOX));
          tma_load_4d(sr, &tm_r, &full_bar[s], 0, p.r_s * j0 + p.r_ox, p.r_s * i0 + p.r_oy, img);
          if (p.vstack)        // a producer pays ~500 cycles per TMA instruction: one 36 KB box instead of nine 4 KB ones
            tma_load_4d(sr + THIN_BOX, &tm_c, &full_bar[s], 0, j0 + p.dx[0], i0 + p.dy[0], img);
          else
            for (int t = 0; t < p.ntaps; ++t)
              tma_load_4d(sr + (1 + t) * THIN_BOX, &tm_c, &full_bar[s], 0, p.c_s * j0 + p.dx[t], p.c_s * i0 + p.dy[t], img);
        }
      }
      __syncwarp();
    }
    }
  } else if (warp == 1) {
    int s = 0; unsigned ph = 0;
    bool first = true;
    for (long long c = cbeg; c < cend; ++c) {
      mbar_wait(&full_bar[s], ph);
      tc_fence_after();
      if (lane == 0) {
        const unsigned r_addr = smem_u32(smem + (size_t)s * p.stage_bytes);
        if (p.merged) {
          // ONE instruction per K step: A = taps 0..3 of the shifted operand (4 column blocks, one row apart), B = the fixed
          // operand at rows i0 - 4*(ngrp-1), .., i0 - 4, i0 (ngrp column blocks, four rows = 16 KB apart):
          //   D[(t1, cc)][(nb, rc)] += sum_p C[p + t1 rows][cc] * R[p - 4*(ngrp-1-nb) rows][rc]   = tap t1 + 4*(ngrp-1-nb)
          // (a chunk is th rows of 64 pixels: both patches hold th - 1 extra rows and K step k is 16 pixels of row k / 4)
          const unsigned c_addr = r_addr + (p.r_rows + p.th - 1) * THIN_BOX;
          for (int k = 0; k < p.th * (THIN_KP / 16); ++k) {
            const unsigned long long ad = make_mn_desc(c_addr + k * 1024, THIN_BOX, 512, 4u);
            const unsigned long long bd = make_mn_desc(r_addr + k * 1024, 4 * THIN_BOX, 512, 4u);
            tc_mma<0>(tmem_base, ad, bd, p.idesc, (first && k == 0) ? 0u : 1u);
          }
        } else
        for (int g = 0; g < p.ngrp; ++g) {
          const unsigned c_addr = r_addr + (1 + 4 * g) * THIN_BOX;
          for (int k = 0; k < THIN_KP / 16; ++k) {           // UMMA_K = 16 pixel rows of 64 bytes
            // A = four tap boxes (column blocks of 32 channels, LBO = one box apart); B = the fixed operand
            const unsigned long long ad = make_mn_desc(c_addr + k * 1024, THIN_BOX, 512, 4u);
            const unsigned long long bd = make_mn_desc(r_addr + k * 1024, THIN_BOX, 512, 4u);
            tc_mma<0>(tmem_base + (unsigned)(g * 32), ad, bd, p.idesc, (first && k == 0) ? 0u : 1u);
          }
        }
        tc_commit(&empty_bar[s]);
        if (c == cend - 1) tc_commit(&tfull_bar);
      }
      first = false;
      __syncwarp();
      if (++s == p.stages) { s = 0; ph ^= 1; }
    }
  } else {
    const int q = warp & 3;                   // TMEM lane quarter = tap within the group; lane = channel of the shifted operand
    mbar_wait(&tfull_bar, 0);
    tc_fence_after();
    for (int g = 0; g < p.ngrp; ++g) {
      float v[32];
      tc_ld32(tmem_base + ((unsigned)(q * 32) << 16) + (unsigned)((p.merged ? p.ngrp - 1 - g : g) * 32), v);
      const int t = 4 * g + q;
      if (t < p.ntaps && lane < p.c_valid) {
        float* o = out + (tap_off ? tap_off[t] : 0) + (long long)lane * p.s_n;
#pragma unroll
        for (int e = 0; e < 32; ++e)
          if (e < p.r_valid) atomicAdd(o + (long long)e * p.s_m, v[e] * p.scale);   // lanes -> consecutive addresses when s_n == 1
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

static int encode_thin_operand(EncodeTiledFn encode, CUtensorMap* tm, const ast_image* im, int tw, int th, int s) {
  cuuint64_t dims[4] = {(cuuint64_t)im->c, (cuuint64_t)im->w, (cuuint64_t)im->h, (cuuint64_t)im->n};
  cuuint64_t strides[3] = {(cuuint64_t)im->sw * 2, (cuuint64_t)im->sh * 2, (cuuint64_t)im->sn * 2};
  cuuint32_t box[4] = {32, (cuuint32_t)(tw * s), (cuuint32_t)(th * s), 1};
  cuuint32_t estr[4] = {1, (cuuint32_t)s, (cuuint32_t)s, 1};
  return cached_tensor_map(encode, tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, im->ptr, dims, strides, box, estr,
                           CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
}

// returns 1 when the thin kernel took the job, 0 when not applicable, < 0 / CUDA error codes otherwise
static int contract_thin(EncodeTiledFn encode, const ast_image* rows, int r_s, int r_oy, int r_ox, const ast_image* cols,
                         int c_s, const short* dy, const short* dx, int ntaps, int mi, int mj, float* out,
                         const int* tap_off, long long s_m, long long s_n, float scale, cudaStream_t stream) {
  if (rows->dtype != AST_BF16 || rows->c > 32 || cols->c > 32 || ntaps < 2 || ntaps > 60) return 0;
  CtThinParams p;
  memset(&p, 0, sizeof(p));
  p.mi = mi; p.mj = mj; p.n_img = rows->n; p.ntaps = ntaps; p.ngrp = (ntaps + 3) / 4;
  p.r_s = r_s; p.r_oy = r_oy; p.r_ox = r_ox; p.c_s = c_s;
  for (int t = 0; t < ntaps; ++t) { p.dy[t] = dy ? dy[t] : 0; p.dx[t] = dx ? dx[t] : 0; }
  p.r_valid = rows->c; p.c_valid = cols->c;
  pick_tile(mi, mj, THIN_KP, &p.tw, &p.th);
  // vertical tap stacks (the 9x9 ends after the row fold: dy = d0 .. d0 + ntaps - 1, one dx, unit coordinate multipliers):
  // with one-row chunks the shifted operand of ALL taps is one box of ntaps rows - 2 TMA instructions per stage, not 10
  p.vstack = c_s == 1 && mj >= THIN_KP / 2;
  for (int t = 1; t < ntaps && p.vstack; ++t) p.vstack = p.dx[t] == p.dx[0] && p.dy[t] == p.dy[0] + t;
  if (p.vstack) { p.tw = THIN_KP; p.th = 1; }
  // merged mode: the three (ntaps = 9) tap groups of a vertical stack become the N blocks of ONE MMA.  The kernel is bound
  // by its MN-major MMAs (M128 x N32 x K16 at ~125 clk each, three per K step); with the fixed operand loaded as a patch
  // of r_rows rows (row shifts 0, -4, -8 as column blocks) tap t1 + 4*t2 is block (t1, t2) of a single M128 x N96 MMA.
  // The chunk rows then run 4*(ngrp-1) rows past the image so that every (row, tap) product is visited.
  p.merged = p.vstack && r_s == 1 && p.ngrp >= 2 && p.ngrp * 32 <= 256;
  p.r_rows = 4 * (p.ngrp - 1) + 1;
  // four image rows per chunk: the row-shifted patches overlap, so the L2 -> shared-memory traffic per image row drops from
  // 9 + 4 to (12 + 7) / 4 patch rows.  First layer (B = 32, 256^2): 172 us unmerged, merged with 1 / 2 / 3 / 4 rows per chunk
  // 131 / 87 / 92 / 80 us (two 76 KB stages at 4 rows).
  if (p.merged) p.th = 4;
  p.tiles_i = (mi + (p.merged ? p.r_rows - 1 : 0) + p.th - 1) / p.th; p.tiles_j = (mj + p.tw - 1) / p.tw;
  p.chunks_total = (long long)p.tiles_i * p.tiles_j * p.n_img;
  p.s_m = s_m; p.s_n = s_n; p.scale = scale;
  p.stage_bytes = p.merged ? (p.r_rows + 4 + 2 * (p.th - 1)) * THIN_BOX : (1 + 4 * p.ngrp) * THIN_BOX;
  p.stages = (200 * 1024) / p.stage_bytes;
  if (p.stages > CT_MAX_STAGES) p.stages = CT_MAX_STAGES;
  ct_pick_producers(p.stages, p.nprod);
  if (p.stages < 2) return 0;
  // bf16 A/B (1), both MN-major (bits 15, 16), N = 32, M = 128
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((unsigned)((p.merged ? p.ngrp * 32 : 32) >> 3) << 17) | ((128u >> 4) << 24);
  long long ks = num_sms();
  if (ks > p.chunks_total) ks = p.chunks_total;
  p.chunks_per_cta = (int)((p.chunks_total + ks - 1) / ks);
  const int grid = (int)((p.chunks_total + p.chunks_per_cta - 1) / p.chunks_per_cta);
  alignas(64) CUtensorMap tm_r, tm_c;
  if (int e = encode_thin_operand(encode, &tm_r, rows, p.tw, p.merged ? p.r_rows + p.th - 1 : p.th, r_s)) return e;
  if (int e = encode_thin_operand(encode, &tm_c, cols, p.tw, p.merged ? 4 + p.th - 1 : (p.vstack ? ntaps : p.th), c_s)) return e;
  const size_t smem = (size_t)p.stages * p.stage_bytes + 1024;
  cudaError_t e = set_max_smem(contract_thin_kernel, smem);
  if (e != cudaSuccess) { set_error("contract_thin: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); return (int)e; }
  launch_k(contract_thin_kernel, grid, CT_THREADS, smem, stream, tm_r, tm_c, p, out, tap_off);
  count_launch();
  count_work(FAM_WGRAD_THIN, 2.0 * rows->n * mi * mj * ntaps * rows->c * cols->c, img_bytes(rows) + img_bytes(cols));
  AST_CUDA_LAUNCH_CHECK();
  return 1;
}

// rows = operand that provides the M (<=128 per block) dimension, cols = the N dimension.
int contract_tc(const ast_image* rows, int r_s, int r_oy, int r_ox, const ast_image* cols, int c_s, const short* dy,
                const short* dx, int ntaps, int mi, int mj, float* out, const int* tap_off, long long s_m,
                long long s_n, long long out_img_stride, float scale, int upper_only, cudaStream_t stream,
                const GramFin* finp) {
  AST_CHECK_ARG(rows->dtype == cols->dtype, "contract_tc: operands must share a dtype");
  AST_CHECK_ARG(rows->n == cols->n, "contract_tc: batch mismatch");
  if (int e = check_operand("contract_tc(rows)", rows)) return e;
  if (int e = check_operand("contract_tc(cols)", cols)) return e;
  AST_CHECK_ARG(r_s >= 1 && r_s <= 2 && c_s >= 1 && c_s <= 2, "contract_tc: coordinate multipliers must be 1 or 2");
  if (rows->n == 0) return 0;
  EncodeTiledFn encode = get_encode();
  AST_CHECK_ARG(encode, "contract_tc: cuTensorMapEncodeTiled entry point not available");
  const int esz = rows->dtype == AST_F32 ? 4 : 2;
  if (!upper_only && out_img_stride == 0 && !finp) {
    const int tr = contract_thin(encode, rows, r_s, r_oy, r_ox, cols, c_s, dy, dx, ntaps, mi, mj, out, tap_off, s_m, s_n,
                                 scale, stream);
    if (tr != 0) return tr == 1 ? 0 : tr;
  }

  CtParams p;
  memset(&p, 0, sizeof(p));
  p.mi = mi; p.mj = mj; p.n_img = rows->n; p.ntaps = ntaps;
  p.r_s = r_s; p.r_oy = r_oy; p.r_ox = r_ox; p.c_s = c_s;
  for (int t = 0; t < ntaps; ++t) { p.dy[t] = dy ? dy[t] : 0; p.dx[t] = dx ? dx[t] : 0; }
  p.cb = 128 / esz;
  p.m_valid = rows->c; p.n_valid = cols->c;
  p.m_boxes = 128 / p.cb;
  const int n_pad = (cols->c + p.cb - 1) / p.cb * p.cb;
  p.bn = n_pad <= 256 ? n_pad : (n_pad % 256 == 0 ? 256 : (n_pad % 128 == 0 ? 128 : p.cb));
  // Gram with the fused finish: 128-wide column blocks when that yields enough single-writer CTAs to fill the GPU without
  // split-K (C = 256 at B = 32: 3 upper blocks x 32 images = 96 CTAs) - the finish then runs from registers
  if (finp && upper_only && n_pad % 128 == 0 && n_pad >= 256) {
    const int mb = (rows->c + 127) / 128, nb = n_pad / 128;
    const long long upper_blocks = (long long)rows->n * (mb * nb - mb * (mb - 1) / 2);
    if (upper_blocks * 10 >= num_sms() * 6) p.bn = 128;
  }
  p.n_boxes = p.bn / p.cb;
  p.m_blocks = (rows->c + 127) / 128;
  p.n_blocks = (n_pad + p.bn - 1) / p.bn;
  p.kp = 64;
  pick_tile(mi, mj, p.kp, &p.tw, &p.th);
  p.tiles_i = (mi + p.th - 1) / p.th; p.tiles_j = (mj + p.tw - 1) / p.tw;
  p.per_img = out_img_stride != 0;
  p.chunks_total = (long long)p.tiles_i * p.tiles_j * (p.per_img ? 1 : p.n_img);
  p.out_img_stride = out_img_stride; p.s_m = s_m; p.s_n = s_n; p.scale = scale; p.upper_only = upper_only;
  p.box_bytes = p.kp * 128;
  // tap groups: one CTA accumulates `tg` taps (tg*bn TMEM columns) from ONE load of the rows operand per stage
  {
    int tg_max = 512 / p.bn;
    if (tg_max > ntaps) tg_max = ntaps;
    while (tg_max > 1 && 2 * (p.m_boxes + tg_max * p.n_boxes) * p.box_bytes > 200 * 1024) --tg_max;
    p.ngroups = (ntaps + tg_max - 1) / tg_max;
    p.tg = (ntaps + p.ngroups - 1) / p.ngroups;       // balance the groups
  }
  p.same = (rows->ptr == cols->ptr && rows->c == cols->c && rows->sn == cols->sn && rows->sh == cols->sh &&
            rows->sw == cols->sw && r_s == c_s && r_oy == 0 && r_ox == 0 && ntaps == 1 && p.dy[0] == 0 && p.dx[0] == 0) ? 1 : 0;
  p.stage_bytes = (p.m_boxes + p.tg * p.n_boxes) * p.box_bytes;
  p.stages = (200 * 1024) / p.stage_bytes;
  if (p.stages > CT_MAX_STAGES) p.stages = CT_MAX_STAGES;
  ct_pick_producers(p.stages, p.nprod);
  AST_CHECK_ARG(p.stages >= 2, "contract_tc: tile does not fit shared memory");
  p.umma_k_bytes = (32 / esz) * 128;                 // UMMA_K pixel rows (16 bf16 / 8 tf32) x 128 B
  p.kmma = p.kp / (32 / esz);
  p.layout_type = esz == 4 ? 1u : 2u;
  p.sbo = esz == 4 ? 512u : 1024u;
  const unsigned fmt = rows->dtype == AST_F32 ? 2u : 1u;
  p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 15) | (1u << 16) | ((unsigned)(p.bn >> 3) << 17) | ((128u >> 4) << 24);
  const long long fixed = (long long)p.ngroups * p.m_blocks * p.n_blocks * (p.per_img ? p.n_img : 1);
  // one CTA per SM (200 KB of smem each): size the split-K so the whole grid is a single wave without a tail
  long long ks = num_sms() / fixed;
  if (finp && fixed * 10 >= num_sms() * 6) ks = 1;     // enough single-writer CTAs: the register finish beats split-K + ticket
  if (ks < 1) ks = 1;
  if (ks > p.chunks_total) ks = p.chunks_total;
  p.chunks_per_cta = (int)((p.chunks_total + ks - 1) / ks);
  p.ksplit = (int)((p.chunks_total + p.chunks_per_cta - 1) / p.chunks_per_cta);

  alignas(64) CUtensorMap tm_r, tm_c;
  if (int e = encode_operand(encode, &tm_r, rows, p.cb, p.tw, p.th, r_s)) return e;
  if (int e = encode_operand(encode, &tm_c, cols, p.cb, p.tw, p.th, c_s)) return e;
  // + m_boxes: a shared A operand (Gram) may read up to m_boxes boxes past the column boxes of the last stage
  const size_t smem = (size_t)p.stages * p.stage_bytes + 1024 + (p.same ? (size_t)p.m_boxes * p.box_bytes : 0);
  dim3 grid((unsigned)(p.ksplit * (p.per_img ? p.n_img : 1)), p.ngroups, p.m_blocks * p.n_blocks);
  GramFin fin;
  memset(&fin, 0, sizeof(fin));
  if (finp) {
    AST_CHECK_ARG(p.per_img && upper_only && s_n == 1 && s_m == cols->c && p.ngroups == 1,
                  "contract_tc: the Gram finishing step needs a per-image upper-triangle launch");
    AST_CHECK_ARG(p.stages * p.stage_bytes >= 4 * 2 * 32 * 33 * (int)sizeof(float), "contract_tc: no room for the mirror tiles");
    AST_CHECK_ARG(p.m_blocks * p.n_blocks <= AST_GRAM_COUNTERS_PER_IMAGE, "contract_tc: C=%d needs more ticket counters", cols->c);
    fin = *finp;
    fin.enabled = 1;
  }
  cudaError_t e;
  if (rows->dtype == AST_BF16) {
    e = set_max_smem(contract_tc_kernel<0>, smem);
    if (e == cudaSuccess) launch_k(contract_tc_kernel<0>, grid, CT_THREADS, smem, stream, tm_r, tm_c, p, out, tap_off, fin);
  } else {
    e = set_max_smem(contract_tc_kernel<1>, smem);
    if (e == cudaSuccess) launch_k(contract_tc_kernel<1>, grid, CT_THREADS, smem, stream, tm_r, tm_c, p, out, tap_off, fin);
  }
  if (e != cudaSuccess) { set_error("contract_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); return (int)e; }
  count_launch();
  if (upper_only)   // Gram: upper-triangle flops C(C+1)HW per image, one read of F + one write of G
    count_work(FAM_GRAM_TC, (double)rows->n * rows->c * (rows->c + 1.0) * mi * mj, img_bytes(rows) + 4.0 * rows->n * rows->c * rows->c);
  else
    count_work(FAM_WGRAD_TC, 2.0 * rows->n * mi * mj * ntaps * rows->c * cols->c, img_bytes(rows) + img_bytes(cols));
  AST_CUDA_LAUNCH_CHECK();
  return 0;
}

// Gram (+ optional fused style-MSE / gradient seed).  With `counters` (zeroed by the caller, AST_GRAM_COUNTERS_PER_IMAGE
// ints per image, like g itself) the finishing CTA of every block mirrors / reduces; without, g is zeroed here and a
// separate kernel mirrors the upper triangle.
int gram_tc(const ast_image* x, float* g, float scale, const float* target, long long target_img_stride, double* loss,
            float loss_scale, float* dmat, float d_scale, int* counters, cudaStream_t s) {
  if (!counters) {
    cudaMemsetAsync(g, 0, sizeof(float) * (size_t)x->n * x->c * x->c, s);
    int rc = contract_tc(x, 1, 0, 0, x, 1, nullptr, nullptr, 1, x->h, x->w, g, nullptr, x->c, 1, (long long)x->c * x->c,
                         scale, 1, s, nullptr);
    if (rc) return rc;
    const long long total = (long long)x->n * x->c * x->c;
    long long blocks = (total + 255) / 256;
    if (blocks > num_sms() * 8) blocks = num_sms() * 8;
    launch_k(mirror_upper_kernel, (int)blocks, 256, 0, s, g, x->c, total);
    count_launch();
    AST_CUDA_LAUNCH_CHECK();
    return 0;
  }
  GramFin fin;
  memset(&fin, 0, sizeof(fin));
  fin.counters = counters; fin.target = target; fin.target_img_stride = target_img_stride;
  fin.loss = loss; fin.loss_scale = loss_scale; fin.dmat = dmat; fin.d_scale = d_scale;
  return contract_tc(x, 1, 0, 0, x, 1, nullptr, nullptr, 1, x->h, x->w, g, nullptr, x->c, 1, (long long)x->c * x->c,
                     scale, 1, s, &fin);
}

}  // namespace ast
