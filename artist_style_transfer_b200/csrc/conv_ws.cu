// Weight-stationary, halo-reusing variant of the tcgen05 gather convolution (stride-1 input, small filters).
//
// conv_tc.cu re-loads the 128-pixel A tile for every tap (9x / 81x read amplification out of L2) and re-streams the
// weights for every tile.  For layers whose whole packed filter fits in shared memory this kernel instead
//   * loads ALL taps' weights once per CTA (they stay resident for every tile the persistent CTA processes), and
//   * loads ONE halo patch (16+kh-1) x 16 pixels per cin-chunk and tile; each tap's A operand is the same patch
//     addressed through a shifted shared-memory descriptor: tile = 16 rows x 8 pixels, one 8-pixel row = one
//     swizzle atom, atoms are SBO = 16 pixels apart, the tap shift (dy*16 + dx) pixels moves the start address
//     (the hardware swizzles on absolute smem address bits, so unaligned starts need no base_offset - measured).
// Used in the training step for: VGG conv1_2 (64 -> 64, fp16 in, fp32 tap + fused MaxPool out) and its data gradient, and the
// 1x1 layers; any stride-1 layer whose filter fits in shared memory is accepted (the thin 32/64-channel layers that ran
// here in round 1 now go through the block-stacked kernel, conv_st.cu, unless a caller asks for the plain launches).
#include "px_common.cuh"

namespace ast {

constexpr int WS_THREADS = 320;   // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (2 per TMEM lane quarter)
constexpr int WS_MAX_PBUF = 4;
constexpr int WS_TH = 16, WS_TW = 8;

struct WsParams {
  int mi, mj, tiles_i, tiles_j, n_img;
  int bn, cout, cout_valid, thin, flags;
  int ntaps, kchunks, kc, rowb;
  int so, oy0, ox0;
  int dy_min, dx_min, ph, pw;
  int patch_bytes, patch_tx, n_pbuf, w_tile_bytes, w_total_bytes;
  int T;                   // sub-tiles (16 rows x 8 px each, stacked vertically) per pipeline step / TMEM stage
  int stage_w;             // bytes of the per-warp store-transpose stage: 1 KB, or 4 KB (a whole 32 x 32 fp32 block) when pooling
  unsigned idesc, layout_type, sbo;
  long long total_tiles;
  short tdy[AST_MAX_TAPS];
  short tdx[AST_MAX_TAPS];
};

__device__ __forceinline__ unsigned long long pack_desc(unsigned lo, unsigned hi) {
  unsigned long long d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}

__device__ __forceinline__ bool valid_window(int iw, int jw, const WsParams& p) { return iw + 1 < p.mi && jw + 1 < p.mj; }
// window code of (a, b, c, d) = positions (0,0), (0,1), (1,0), (1,1): bits 0-1 arg max (first maximum wins, like ATen),
// bits 2-5 = (value > 0); m = the maximum (same definition as ast_maxpool2_fwd / pool_code in pointwise.cu)
__device__ __forceinline__ unsigned pool_code4(float a, float b, float c, float d, float& m) {
  int arg = 0; m = a;
  if (b > m) { m = b; arg = 1; }
  if (c > m) { m = c; arg = 2; }
  if (d > m) { m = d; arg = 3; }
  return (unsigned)arg | ((a > 0.f) ? 4u : 0u) | ((b > 0.f) ? 8u : 0u) | ((c > 0.f) ? 16u : 0u) | ((d > 0.f) ? 32u : 0u);
}

// Epilogue of one 32-pixel x 32-channel chunk WITH the fused nn.MaxPool2d(2, 2) (+ window codes): bias / ReLU / TF32
// rounding in registers, then the whole fp32 block goes through a 4 KB per-warp shared-memory stage (pixel-major,
// XOR-swizzled 16-byte chunks).  From the stage (a) the full-resolution output is written with coalesced 16-byte stores
// and (b) every lane pools ONE 2 x 2 window x 8 channels with plain ALU code: lane = window (0..7) x channel group (0..3).
// (The first version exchanged the window's values between lanes: 3 shuffles + a compare chain per channel in EVERY
// lane, ~480 of the ~800 instructions of the chunk - conv1_2 was bound by it once fp16 operands had halved its MMAs.)
__device__ __forceinline__ void ws_epilogue_pooled(float* v, int co, int img, int i_base, int j_base, bool valid, bool store_out,
                                                   const WsParams& p, const float* __restrict__ bias, const Img& out,
                                                   const EpiRows& rows, const Img& pooled, const Img& pcodes,
                                                   unsigned char* stage, int lane) {
  if (co >= p.cout) return;                                  // uniform
  if (bias) {
#pragma unroll
    for (int e = 0; e < 32; e += 4) {
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + co + e));
      v[e] += b4.x; v[e + 1] += b4.y; v[e + 2] += b4.z; v[e + 3] += b4.w;
    }
  }
  if (p.flags & AST_CONV_RELU) {
#pragma unroll
    for (int e = 0; e < 32; ++e) v[e] = fmaxf(v[e], 0.f);
  }
  if (p.flags & AST_CONV_ROUND_TF32) {
#pragma unroll
    for (int e = 0; e < 32; ++e) v[e] = round_tf32(v[e]);
  }
  const unsigned st = smem_u32(stage);
  // pixel `lane` -> stage row `lane` (128 B), chunk c at 16 * (c ^ (lane & 7))
#pragma unroll
  for (int c = 0; c < 8; ++c)
    sts128(st + 128u * lane + 16u * (c ^ (lane & 7)),
           make_uint4(__float_as_uint(v[4 * c]), __float_as_uint(v[4 * c + 1]), __float_as_uint(v[4 * c + 2]), __float_as_uint(v[4 * c + 3])));
  __syncwarp();
  if (store_out && out.dtype == AST_F32) {                   // pixel 4k + lane/8, chunk lane%8: 128-byte runs per pixel
    float* base = (float*)out.ptr + co + 4 * (lane & 7);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int px = 4 * k + (lane >> 3);
      const uint4 val = lds128(st + 128u * px + 16u * ((lane & 7) ^ (px & 7)));
      const long long off = rows.off[k];
      if (off >= 0) *reinterpret_cast<uint4*>(base + off) = val;
    }
  }
  // ---- pooling: lane = (window wi = lane / 4, channels 8 * (lane % 4) ..); the warp's pixels are 4 tile rows x 8 columns
  const int wi = lane >> 2, cg = lane & 3;
  const int wy = wi >> 2, wx = wi & 3;
  const int iw = i_base + 2 * wy, jw = j_base + 2 * wx;
  float a[4][8];
#pragma unroll
  for (int k = 0; k < 4; ++k) {                              // k = 2 * dy + dx
    const int px = 8 * (2 * wy + (k >> 1)) + 2 * wx + (k & 1);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const uint4 u = lds128(st + 128u * px + 16u * ((2 * cg + h) ^ (px & 7)));
      a[k][4 * h] = __uint_as_float(u.x); a[k][4 * h + 1] = __uint_as_float(u.y);
      a[k][4 * h + 2] = __uint_as_float(u.z); a[k][4 * h + 3] = __uint_as_float(u.w);
    }
  }
  __syncwarp();                                              // the stage is rewritten by the next chunk
  if (!valid_window(iw, jw, p)) return;
  float m[8];
  unsigned c0 = 0, c1 = 0;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const unsigned code = pool_code4(a[0][e], a[1][e], a[2][e], a[3][e], m[e]);
    if (e < 4) c0 |= code << (8 * e); else c1 |= code << (8 * (e - 4));
  }
  const long long po = img_off(pooled, img, iw >> 1, jw >> 1, co + 8 * cg);
  st4_img(pooled, po, m);
  st4_img(pooled, po + 4, m + 4);
  if (pcodes.ptr) *reinterpret_cast<uint2*>(pcodes.ptr + img_off(pcodes, img, iw >> 1, jw >> 1, co + 8 * cg)) = make_uint2(c0, c1);
}

// MINB = 2: built for two resident CTAs per SM (<= 102 registers): layers whose weights + patches need less than half of
// the shared memory (the 32-channel 256^2 layers, conv1_1) are latency bound with one persistent CTA per SM.
template <int KIND, int MINB>
__global__ void __launch_bounds__(WS_THREADS, MINB)
conv_ws_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_w, const WsParams p,
               const float* __restrict__ bias, const Img add, const Img mask, const Img out,
               double* __restrict__ stats, const Img pooled, const Img pcodes) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long pfull[WS_MAX_PBUF], pempty[WS_MAX_PBUF], tfull_bar[2], tempty_bar[2], wbar;
  __shared__ unsigned tmem_slot;
  __shared__ unsigned s_tapoff[AST_MAX_TAPS];    // per-tap A start offset inside the patch, in 16-byte units

  unsigned char* smem_w = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char* smem_p = smem_w + ((p.w_total_bytes + 1023) & ~1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int acc_cols = p.T * p.bn;               // one TMEM stage: T accumulators of bn columns
  const unsigned tmem_cols = (2 * acc_cols <= 32) ? 32 : (2 * acc_cols <= 64) ? 64 : (2 * acc_cols <= 128) ? 128 : (2 * acc_cols <= 256) ? 256 : 512;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_in) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_w) : "memory");
    for (int s = 0; s < p.n_pbuf; ++s) { mbar_init(&pfull[s], 1); mbar_init(&pempty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 256); }
    mbar_init(&wbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (threadIdx.x >= 64 && threadIdx.x - 64 < p.ntaps) {
    const int t = threadIdx.x - 64;
    s_tapoff[t] = (unsigned)((p.tdy[t] * p.pw + p.tdx[t]) * p.rowb) >> 4;
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem_base = tmem_slot;

  if (warp == 0) {
    // ============================ TMA producer ============================
    if (lane == 0) {       // the whole filter, once
      mbar_expect_tx(&wbar, (unsigned)p.w_total_bytes);
      for (int t = 0; t < p.ntaps; ++t)
        for (int kc = 0; kc < p.kchunks; ++kc)
          tma_load_2d(smem_w + (size_t)(t * p.kchunks + kc) * p.w_tile_bytes, &tm_w, &wbar, kc * p.kc, t * p.cout);
    }
    __syncwarp();
    int s = 0; unsigned ph = 0;
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      long long r = tile;
      const int tj = (int)(r % p.tiles_j); r /= p.tiles_j;
      const int ti = (int)(r % p.tiles_i);
      const int img = (int)(r / p.tiles_i);
      const int x0 = tj * WS_TW + p.dx_min, y0 = ti * WS_TH * p.T + p.dy_min;
      for (int kc = 0; kc < p.kchunks; ++kc) {
        mbar_wait(&pempty[s], ph ^ 1);
        if (lane == 0) {
          mbar_expect_tx(&pfull[s], (unsigned)p.patch_tx);
          tma_load_4d(smem_p + (size_t)s * p.patch_bytes, &tm_in, &pfull[s], kc * p.kc, x0, y0, img);
        }
        __syncwarp();
        if (++s == p.n_pbuf) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    int s = 0; unsigned ph = 0; int as = 0; unsigned aph = 0;
    const int kmma = p.rowb / 32;
    mbar_wait(&wbar, 0);
    tc_fence_after();
    const unsigned w_addr0 = smem_u32(smem_w);
    const unsigned hi_a = (p.sbo >> 4) | (1u << 14) | (p.layout_type << 29);
    const unsigned hi_b = ((8u * p.rowb) >> 4) | (1u << 14) | (p.layout_type << 29);
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      mbar_wait(&tempty_bar[as], aph ^ 1);
      tc_fence_after();
      const unsigned d_tmem0 = tmem_base + (unsigned)(as * acc_cols);
      for (int kc = 0; kc < p.kchunks; ++kc) {
        mbar_wait(&pfull[s], ph);
        tc_fence_after();
        if (lane == 0) {
          // descriptors: hi word is constant, lo word = (addr >> 4) | LBO; one 32-bit add per MMA
          const unsigned p_lo = ((smem_u32(smem_p + (size_t)s * p.patch_bytes) & 0x3FFFFu) >> 4) | (1u << 16);
          const unsigned w_lo = ((w_addr0 & 0x3FFFFu) >> 4) | (1u << 16);
          const unsigned w16 = (unsigned)p.w_tile_bytes >> 4;
          const unsigned sub16 = (unsigned)(WS_TH * p.pw * p.rowb) >> 4;      // patch offset of the next sub-tile (16 rows down)
#pragma unroll 1
          for (int st = 0; st < p.T; ++st) {
            const unsigned d_tmem = d_tmem0 + (unsigned)(st * p.bn);
            unsigned acc = kc > 0 ? 1u : 0u;
            const unsigned pa = p_lo + (unsigned)st * sub16;
            unsigned b_lo = w_lo + (unsigned)kc * w16;
            const unsigned b_step = (unsigned)p.kchunks * w16;
            // the issuing thread must stay ahead of the tensor pipe (65-128 clk per instruction): straight-line code,
            // tap offsets prefetched three at a time, no multiplies in the loop
#pragma unroll 3
            for (int t = 0; t < p.ntaps; ++t) {
              const unsigned a_lo = pa + s_tapoff[t];
              tc_mma<KIND>(d_tmem, pack_desc(a_lo, hi_a), pack_desc(b_lo, hi_b), p.idesc, acc);
              tc_mma<KIND>(d_tmem, pack_desc(a_lo + 2, hi_a), pack_desc(b_lo + 2, hi_b), p.idesc, 1u);
              if (kmma == 4) {
                tc_mma<KIND>(d_tmem, pack_desc(a_lo + 4, hi_a), pack_desc(b_lo + 4, hi_b), p.idesc, 1u);
                tc_mma<KIND>(d_tmem, pack_desc(a_lo + 6, hi_a), pack_desc(b_lo + 6, hi_b), p.idesc, 1u);
              }
              acc = 1u;
              b_lo += b_step;
            }
          }
          tc_commit(&pempty[s]);
          if (kc == p.kchunks - 1) tc_commit(&tfull_bar[as]);
        }
        __syncwarp();
        if (++s == p.n_pbuf) { s = 0; ph ^= 1; }
      }
      if (++as == 2) { as = 0; aph ^= 1; }
    }
  } else {
    // ============================ epilogue ============================
    const int q = warp & 3;                        // TMEM lane quarter of this warp
    const int cpar = (warp - 2) >> 2;              // which half of the 32-column chunks this warp drains
    const int row = q * 32 + lane;
    const int ty = row / WS_TW, tx = row % WS_TW;
    unsigned char* stage = smem_p + (size_t)p.n_pbuf * p.patch_bytes + (warp - 2) * p.stage_w;
    int as = 0; unsigned aph = 0;
    EpiStatAcc sacc;
    sacc.reset(-1);
    const int s_chunks = (p.bn - cpar * 32 + 63) / 64;          // 32-channel chunks this warp drains per tile
    const bool s_run = stats && s_chunks <= 2;                  // running sums (else: immediate atomics)
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      long long r = tile;
      const int tj = (int)(r % p.tiles_j); r /= p.tiles_j;
      const int ti = (int)(r % p.tiles_i);
      const int img = (int)(r / p.tiles_i);
      if (s_run && img != sacc.img) { sacc.flush(stats, p.cout_valid, cpar * 32, lane, s_chunks); sacc.reset(img); }
      for (int st = 0; st < p.T; ++st) {
        const int i = (ti * p.T + st) * WS_TH + ty, j = tj * WS_TW + tx;
        const int oy = p.oy0 + p.so * i, ox = p.ox0 + p.so * j;
        const bool valid = i < p.mi && j < p.mj && oy < out.h && ox < out.w;
        float pm[32];                                   // mask of this warp's first chunk, loaded while the MMAs still run
        const bool use_pm = MINB == 1 && mask.ptr && !p.thin && cpar * 32 < p.bn;   // (the 2-CTA build has no registers to spare)
        if (use_pm) tc_epi_prefetch_mask(mask, img, oy, ox, cpar * 32, valid, pm);
        if (st == 0) { mbar_wait(&tfull_bar[as], aph); tc_fence_after(); }
        const unsigned taddr0 = tmem_base + ((unsigned)(q * 32) << 16) + (unsigned)(as * acc_cols + st * p.bn);
        EpiRows rows;
        const bool store_out = !(p.flags & AST_CONV_POOL_ONLY);
        if (!p.thin) tc_epi_row_offsets((valid && store_out) ? img_off(out, img, oy, ox, 0) : -1, lane, out.dtype == AST_F32, rows);
        for (int c0 = cpar * 32; c0 < p.bn; c0 += 64) {
          float v[32];
          tc_ld32(taddr0 + c0, v);
          const int co = c0;
          if (p.thin) {
            if (valid) tc_epilogue32(v, co, img, oy, ox, true, p.cout, p.cout_valid, p.flags, bias, add, mask, out);
          } else {
            if (stats) {
              float su, sq;
              tc_epi_stats_reduce(v, valid, lane, su, sq);
              if (s_run) { const int k = (c0 - cpar * 32) >> 6; sacc.s[k & 1][0] += (double)su; sacc.s[k & 1][1] += (double)sq; }
              else { double* row = stats + ((long long)img * p.cout_valid + co + lane) * 2; atomicAdd(row, (double)su); atomicAdd(row + 1, (double)sq); }
            }
            if (MINB == 1 && pooled.ptr)           // (pooled launches never use the two-CTA build: 32 KB of stage memory)
              ws_epilogue_pooled(v, co, img, (ti * p.T + st) * WS_TH + 4 * q, tj * WS_TW, valid, store_out, p, bias, out, rows,
                                 pooled, pcodes, stage, lane);
            else
              tc_epilogue32_coalesced(v, co, img, oy, ox, valid, p.cout, p.flags, bias, add, mask, out, rows, stage, lane,
                                      pm, use_pm && c0 == cpar * 32);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[as]);
      if (++as == 2) { as = 0; aph ^= 1; }
    }
    if (s_run) sacc.flush(stats, p.cout_valid, cpar * 32, lane, s_chunks);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// Returns 1 = launched, 0 = not applicable (caller uses conv_px / conv_tc), other = error.
// Applies when the input stride is 1, the taps span <= 9 x 16 pixels and the whole packed filter plus two halo patches fit
// in shared memory (VGG conv1_x and their dgrads, the 32-channel 256^2 layers, the vertical-tap forms of the 9x9 layers).
// Variants measured slower and removed (scratch/dead_variants/conv_ws_r01.cu.txt, profiles/r01_summary.md finding 11):
// streamed weights, a pixels-as-N form with 16 epilogue warps, base_offset descriptors.
int conv_gather_ws(const ast_image* in, const void* weights, const float* bias, const ast_image* add,
                   const ast_image* mask, const ast_image* out, const ast_gather_geom* g, int cpad, bool thin,
                   cudaStream_t stream) {
  if (g->si != 1 || g->w_img_stride != 0) return 0;
  if (g->pooled) {          // fused MaxPool2d(2,2) output: plain stride-1 launches with a full NHWC output only
    const ast_image* q = g->pooled;
    if (thin || g->so != 1 || g->oy0 != 0 || g->ox0 != 0) return 0;
    AST_CHECK_ARG(q->n == in->n && q->h == g->mi / 2 && q->w == g->mj / 2 && q->c == out->c && q->sc == 1 &&
                  q->sw % 4 == 0 && q->sh % 4 == 0 && q->sn % 4 == 0 && ((uintptr_t)q->ptr & 15) == 0,
                  "conv_ws: pooled must be [n, mi/2, mj/2, cout] NHWC with 16-byte aligned pixels");
    AST_CHECK_ARG(out->dtype == AST_F32 || (g->flags & AST_CONV_POOL_ONLY), "conv_ws: the pooled launch stores an fp32 full-resolution output (or none: AST_CONV_POOL_ONLY)");
    AST_CHECK_ARG(!add && !mask, "conv_ws: the pooled launch takes no add / mask operand");
    const ast_image* pc = g->pool_codes;
    AST_CHECK_ARG(!pc || (pc->dtype == AST_U8 && same_shape(pc, q) && pc->sc == 1 && pc->sw % 16 == 0 && pc->sh % 16 == 0 &&
                          pc->sn % 16 == 0 && ((uintptr_t)pc->ptr & 15) == 0 && pc->c % 32 == 0),
                  "conv_ws: pool_codes must be uint8 [n, mi/2, mj/2, cout] NHWC with 16-byte aligned pixels");
  }
  const int esz = in->dtype == AST_F32 ? 4 : 2;
  int dy_min = 1 << 30, dy_max = -(1 << 30), dx_min = 1 << 30, dx_max = -(1 << 30);
  for (int t = 0; t < g->ntaps; ++t) {
    dy_min = g->dy[t] < dy_min ? g->dy[t] : dy_min; dy_max = g->dy[t] > dy_max ? g->dy[t] : dy_max;
    dx_min = g->dx[t] < dx_min ? g->dx[t] : dx_min; dx_max = g->dx[t] > dx_max ? g->dx[t] : dx_max;
  }
  if (dx_max - dx_min > 8 || dy_max - dy_min > 15) return 0;
  if (cpad > 256) return 0;
  // multi-tap layers with >= 128 output channels belong to the halo pixels-as-N kernel (conv_hx.cu) even when their 16-bit
  // filter would fit here (VGG conv2_1 with fp16 operands: 146 vs 122 us)
  if (cpad % 128 == 0 && g->ntaps >= 8 && !thin && !g->pooled && (in->c * esz) % 128 == 0) return 0;
  WsParams p;
  memset(&p, 0, sizeof(p));
  p.rowb = (in->c * esz) % 128 == 0 ? 128 : 64;
  p.kc = p.rowb / esz;
  p.kchunks = in->c / p.kc;
  p.bn = cpad;
  p.w_tile_bytes = p.bn * p.rowb;
  p.w_total_bytes = g->ntaps * p.kchunks * p.w_tile_bytes;
  p.pw = (dx_max == dx_min) ? WS_TW : 16;
  p.stage_w = g->pooled ? 4096 : 1024;
  const int stage_total = 8 * p.stage_w;
  const int avail = 232448 - 1024 - 1024 - stage_total;  // align slack, static smem, store-transpose stage
  const int budget = avail - ((p.w_total_bytes + 1023) & ~1023);
  // Sub-tiles per pipeline step: T vertically stacked 16 x 8 sub-tiles share one patch load (less halo), one TMEM stage
  // (T x bn columns) and one round of the producer -> MMA -> epilogue barrier chain.  Measured (B=32, 256^2 step): the
  // 4-phase ConvTranspose 64->32 337 -> 252 us with T = 4, 128->64 133 -> 107 us with T = 2; no gain for cout >= 128.
  // (What bounds the 32-channel layers is the 64-byte pixel row: both the TMA patch load and the tensor core's operand
  // fetch move 64-byte rows at the cost of 128-byte ones - measured 160 clk per M128 x N32 MMA.)
  p.T = 1;
  const int t_max = p.bn <= 32 ? 4 : (p.bn <= 64 ? 2 : 1);
  for (int T = t_max; T >= 1; --T) {
    if (2 * T * p.bn > 512 || (T > 1 && (T - 1) * WS_TH >= g->mi)) continue;
    const int pb = ((p.pw * (WS_TH * T + (dy_max - dy_min)) * p.rowb) + 1023) & ~1023;
    if (budget >= (T > 1 ? 3 : 2) * pb) { p.T = T; break; }
  }
  p.ph = WS_TH * p.T + (dy_max - dy_min);
  p.patch_tx = p.pw * p.ph * p.rowb;
  p.patch_bytes = (p.patch_tx + 1023) & ~1023;
  if (budget < 2 * p.patch_bytes) return 0;              // the filter does not fit next to two patches
  p.n_pbuf = budget / p.patch_bytes;
  if (p.n_pbuf > WS_MAX_PBUF) p.n_pbuf = WS_MAX_PBUF;
  EncodeTiledFn encode = get_encode();
  AST_CHECK_ARG(encode, "conv_ws: cuTensorMapEncodeTiled entry point not available");

  p.mi = g->mi; p.mj = g->mj; p.so = g->so; p.oy0 = g->oy0; p.ox0 = g->ox0;
  p.ntaps = g->ntaps; p.flags = g->flags; p.cout = cpad; p.cout_valid = out->c; p.thin = thin; p.n_img = in->n;
  p.dy_min = dy_min; p.dx_min = dx_min;
  for (int t = 0; t < g->ntaps; ++t) { p.tdy[t] = g->dy[t] - dy_min; p.tdx[t] = g->dx[t] - dx_min; }
  p.tiles_i = (p.mi + WS_TH * p.T - 1) / (WS_TH * p.T);
  p.tiles_j = (p.mj + WS_TW - 1) / WS_TW;
  p.total_tiles = (long long)p.n_img * p.tiles_i * p.tiles_j;
  p.layout_type = p.rowb == 128 ? 2u : 4u;
  p.sbo = (unsigned)(p.pw * p.rowb);
  // Measured on B200: the tensor core applies the 128B/64B swizzle XOR on ABSOLUTE shared-memory address bits, so a
  // start address shifted by whole pixels needs base_offset = 0 (setting it to (addr>>7)&7 gives wrong results).
  const unsigned fmt = tc_operand_fmt(in->dtype);
  p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((unsigned)(p.bn >> 3) << 17) | ((128u >> 4) << 24);

  const CUtensorMapDataType dt = tc_tmap_dtype(in->dtype);
  const CUtensorMapSwizzle sw = p.rowb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  alignas(64) CUtensorMap tm_in, tm_w;
  {
    cuuint64_t dims[4] = {(cuuint64_t)in->c, (cuuint64_t)in->w, (cuuint64_t)in->h, (cuuint64_t)in->n};
    cuuint64_t strides[3] = {(cuuint64_t)in->sw * esz, (cuuint64_t)in->sh * esz, (cuuint64_t)in->sn * esz};
    cuuint32_t box[4] = {(cuuint32_t)p.kc, (cuuint32_t)p.pw, (cuuint32_t)p.ph, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    if (int r = cached_tensor_map(encode, &tm_in, dt, 4, in->ptr, dims, strides, box, estr, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B)) return r;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)in->c, (cuuint64_t)((long long)g->ntaps * cpad)};
    cuuint64_t strides[1] = {(cuuint64_t)in->c * esz};
    cuuint32_t box[2] = {(cuuint32_t)p.kc, (cuuint32_t)p.bn};
    cuuint32_t estr[2] = {1, 1};
    if (int r = cached_tensor_map(encode, &tm_w, dt, 2, const_cast<void*>(weights), dims, strides, box, estr, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B)) return r;
  }
  // two CTAs per SM when both fit: layers whose weights + 3 patches need less than half of the shared memory (the
  // 32-channel 256^2 layers, conv1_1) are latency bound with one persistent CTA per SM
  bool two = false;
  {
    const int nb = p.n_pbuf > 3 ? 3 : p.n_pbuf;
    const size_t need = 1024 + ((p.w_total_bytes + 1023) & ~1023) + (size_t)nb * p.patch_bytes + stage_total;
    const int acc2 = 2 * p.T * p.bn;                    // TMEM columns one CTA allocates (rounded up to a power of two)
    const int cols = acc2 <= 32 ? 32 : acc2 <= 64 ? 64 : acc2 <= 128 ? 128 : acc2 <= 256 ? 256 : 512;
    if (!g->pooled && 2 * (need + 1024) <= 227 * 1024 && 2 * cols <= 512) { two = true; p.n_pbuf = nb; }
  }
  const size_t smem = 1024 + ((p.w_total_bytes + 1023) & ~1023) + (size_t)p.n_pbuf * p.patch_bytes + stage_total;
  const int max_ctas = (two ? 2 : 1) * num_sms();
  const int grid = (int)(p.total_tiles < max_ctas ? p.total_tiles : max_ctas);
  Img addi = add ? to_img(add) : null_img(), maski = mask ? to_img(mask) : null_img();
  Img pooli = g->pooled ? to_img(g->pooled) : null_img();
  Img codei = (g->pooled && g->pool_codes) ? to_img(g->pool_codes) : null_img();
  cudaError_t e;
#define WS_LAUNCH(K, B)                                                                                          \
  e = set_max_smem(conv_ws_kernel<K, B>, smem);                                                                   \
  if (e == cudaSuccess) launch_k(conv_ws_kernel<K, B>, grid, WS_THREADS, smem, stream, tm_in, tm_w, p, bias, addi, maski, to_img(out), g->stats, pooli, codei)
  if (in->dtype != AST_F32) {          // kind::f16 (bf16 or fp16 operands, the format is in the instruction descriptor)
    if (two) { WS_LAUNCH(0, 2); } else { WS_LAUNCH(0, 1); }
  } else {
    if (two) { WS_LAUNCH(1, 2); } else { WS_LAUNCH(1, 1); }
  }
#undef WS_LAUNCH
  if (e != cudaSuccess) { set_error("conv_ws: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); return (int)e; }
  count_launch();
  count_work(FAM_CONV_WS, conv_flops(in, out, g), conv_bytes(in, out, g, add, mask));
  AST_CUDA_LAUNCH_CHECK();
  return 1;
}

}  // namespace ast
