// Halo-reusing, pixels-as-N tcgen05 gather convolution with STREAMED weights, for stride-1 multi-tap layers whose
// output-channel count is a multiple of 128 (residual 3x3 convs and their data gradients, VGG conv2_x .. conv4_x forward
// and data gradients).
//
// conv_px.cu reloads the pixel operand for every tap: a (tap, cin-chunk) stage moves 16 KB of weights + 32 KB of pixels
// for 4 MMAs (512 cycles), i.e. 94 B/clk per SM against the ~60 B/clk an SM pulls out of L2 - ncu showed the tensor pipe
// 62 % active.  Here ONE halo patch (R+2) x 10 pixels per cin-chunk serves all nine taps through shifted shared-memory
// descriptors (as in conv_ws.cu: the hardware swizzles on absolute address bits, so a start address moved by whole
// 128-byte pixel rows needs no base_offset), and only the weights stream: per 256-pixel tile and cin-chunk 43.5 KB of
// pixels + 9 x 16 KB of weights feed 36 MMAs (4608 cycles) = 41 B/clk.
//
//     D^T[co][pixel] (fp32, TMEM: lane = output channel of the 128-channel slice, column = pixel r*8 + x of an R x 8 tile)
//         += W_t[co][k] * Patch[(r + dy_t) * 10 + x + dx_t][k]
//
// Warp roles: warp 0 = patch TMA producer, warp 10 = weight TMA producer (a separate warp: two divergent spin-wait loops
// inside ONE warp time-slice each other, measured 4x slower), warp 1 = MMA issuer, warps 2-9 = epilogue (one thread = one
// output channel x 32 pixels of a tcgen05.ld, px_common.cuh).
// Measured on B200 (profiles/r02_summary.md):
//  * a producer pays ~470-620 cycles per TMA *instruction* (wait + expect_tx + issue) however many lanes of the warp take
//    turns, so 16 KB instructions cap the feed at ~32 B/clk; the weights therefore come through a 3-D map
//    [tap][cout][cin] with a box of `tg` = 3 taps x 128 rows: ONE 48 KB instruction per slot;
//  * an isolated issue loop retires one M128 x N256 x 32 B MMA per 128 clk (65 clk for N <= 128) from shared memory with
//    a different operand tile every instruction (scratch/mma_bench2.cu) - i.e. the tensor core alone pulls 96-126 B/clk
//    of the 128 B/clk shared-memory bandwidth.  In this kernel, with every TMA wait removed, the same MMAs retire one per
//    ~170-210 clk: the TMA fills and the epilogue compete for the remaining shared-memory bandwidth.  That (not L2) is what
//    bounds conv_px (94 B/clk of TMA fill per stage), this kernel (41 B/clk) and conv_ws; going further needs operands
//    that do not come from this CTA's shared memory (cta_group::2 halves B per SM, or A from TMEM).
#include "px_common.cuh"

namespace ast {

constexpr int HX_EPI_WARPS = 8;                          // 2 per TMEM lane quarter (16, at 96 registers, were measured slower: residual 3x3 44 -> 47 us, masked VGG dgrads up to 2x)
constexpr int HX_WPROD_WARP = 2 + HX_EPI_WARPS;          // weight TMA producer
constexpr int HX_THREADS = 32 * (HX_WPROD_WARP + 1);     // warp 0 patch TMA, warp 1 MMA, then the epilogue warps, then the weight TMA warp
constexpr int HX_MAX_PBUF = 4;
constexpr int HX_MAX_WBUF = 4;
constexpr int HX_TW = 8;

struct HxParams {
  int mi, mj, tiles_i, tiles_j, n_img, n_slices, R;
  int ntaps, kchunks, kc, cout, flags;
  // output mapping: the 128 TMEM lanes of a slice are nblk blocks of cb channels (plain launches: one block of 128);
  // block g writes pixel (oy[g] + soy * i, ox[g] + sox * j) for grid point (i, j), which reads input row sy * i + dy
  int nblk, cb, sy, soy, sox, cstat;                         // cstat: channels per image in the statistics array
  int oy[4], ox[4];
  int dy_min, dx_min, ph, pw;
  int patch_bytes, patch_tx, n_pbuf, n_wbuf, tg, ngroups;   // tg taps per weight slot (one TMA instruction), ngroups per cin-chunk
  unsigned idesc;
  long long total_tiles;
  short tdy[AST_MAX_TAPS];
  short tdx[AST_MAX_TAPS];
};

__device__ __forceinline__ void hx_tile(const HxParams& p, long long tile, int& slice, int& tj, int& ti, int& img) {
  long long r = tile;
  slice = (int)(r % p.n_slices); r /= p.n_slices;
  tj = (int)(r % p.tiles_j); r /= p.tiles_j;
  ti = (int)(r % p.tiles_i);
  img = (int)(r / p.tiles_i);
}

template <int KIND>
__global__ void __launch_bounds__(HX_THREADS, 1)
conv_hx_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_w, const HxParams p,
               const float* __restrict__ bias, const Img32 add, const Img32 mask, const Img32 out,
               double* __restrict__ stats) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long pfull[HX_MAX_PBUF], pempty[HX_MAX_PBUF], wfull[HX_MAX_WBUF], wempty[HX_MAX_WBUF];
  __shared__ __align__(8) unsigned long long tfull_bar[2], tempty_bar[2];
  __shared__ unsigned tmem_slot;
  __shared__ unsigned s_tapoff[AST_MAX_TAPS];    // per-tap start offset of the pixel operand inside the patch, 16-byte units

  unsigned char* smem_p = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char* smem_w = smem_p + (size_t)p.n_pbuf * p.patch_bytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int W_TILE = 128 * 128;              // one (tap, cin-chunk) weight tile: 128 rows x 128 bytes
  const int w_slot = p.tg * W_TILE;              // one ring slot: tg consecutive taps

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_in) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_w) : "memory");
    for (int s = 0; s < p.n_pbuf; ++s) { mbar_init(&pfull[s], 1); mbar_init(&pempty[s], 1); }
    for (int s = 0; s < p.n_wbuf; ++s) { mbar_init(&wfull[s], 1); mbar_init(&wempty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 32 * HX_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (threadIdx.x >= 64 && threadIdx.x - 64 < p.ntaps) {
    const int t = threadIdx.x - 64;
    s_tapoff[t] = (unsigned)((p.tdy[t] * p.pw + p.tdx[t]) * 128) >> 4;
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem_base = tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ============================ patch producer ============================
      int s = 0; unsigned ph = 0;
      for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int slice, tj, ti, img;
        hx_tile(p, tile, slice, tj, ti, img);
        const int x0 = tj * HX_TW + p.dx_min, y0 = ti * p.R * p.sy + p.dy_min;
        for (int kc = 0; kc < p.kchunks; ++kc) {
          mbar_wait(&pempty[s], ph ^ 1);
          mbar_expect_tx(&pfull[s], (unsigned)p.patch_tx);
          tma_load_4d(smem_p + (size_t)s * p.patch_bytes, &tm_in, &pfull[s], kc * p.kc, x0, y0, img);
          if (++s == p.n_pbuf) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == HX_WPROD_WARP) {
    if (lane == 0) {
      // ============================ weight producer: one instruction per (cin-chunk, tap group) ============================
      int s = 0; unsigned ph = 0;
      for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int slice, tj, ti, img;
        hx_tile(p, tile, slice, tj, ti, img);
        for (int kc = 0; kc < p.kchunks; ++kc) {
          for (int gidx = 0; gidx < p.ngroups; ++gidx) {
            mbar_wait(&wempty[s], ph ^ 1);
            mbar_expect_tx(&wfull[s], (unsigned)w_slot);          // taps past the last one are zero-filled: full box bytes
            tma_load_3d(smem_w + (size_t)s * w_slot, &tm_w, &wfull[s], kc * p.kc, slice * 128, gidx * p.tg);
            if (++s == p.n_wbuf) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    int ps = 0; unsigned pph = 0; int ws = 0; unsigned wph = 0; int as = 0; unsigned aph = 0;
    const unsigned hi_w = ((8u * 128u) >> 4) | (1u << 14) | (2u << 29);                 // weights: dense 128-byte rows
    const unsigned hi_p = (((unsigned)(p.sy * p.pw) * 128u) >> 4) | (1u << 14) | (2u << 29);   // patch: 8-pixel tile rows, SBO = sy row pitches
    const unsigned w_base = smem_u32(smem_w), p_base = smem_u32(smem_p);
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      mbar_wait(&tempty_bar[as], aph ^ 1);
      tc_fence_after();
      const unsigned d_tmem = tmem_base + (unsigned)(as * 256);
      for (int kc = 0; kc < p.kchunks; ++kc) {
        mbar_wait(&pfull[ps], pph);
        const unsigned p_lo = (((p_base + (unsigned)ps * (unsigned)p.patch_bytes) & 0x3FFFFu) >> 4) | (1u << 16);
        for (int gidx = 0; gidx < p.ngroups; ++gidx) {
          mbar_wait(&wfull[ws], wph);
          tc_fence_after();
          if (lane == 0) {
            const unsigned w_lo = (((w_base + (unsigned)ws * (unsigned)w_slot) & 0x3FFFFu) >> 4) | (1u << 16);
            const int t0 = gidx * p.tg, t1 = min(p.ntaps, t0 + p.tg);
            for (int t = t0; t < t1; ++t) {
              const unsigned a_lo = w_lo + (unsigned)(t - t0) * (W_TILE >> 4);
              const unsigned b_lo = p_lo + s_tapoff[t];
              tc_mma<KIND>(d_tmem, pack_desc64(a_lo, hi_w), pack_desc64(b_lo, hi_p), p.idesc, (kc | t) ? 1u : 0u);
              tc_mma<KIND>(d_tmem, pack_desc64(a_lo + 2, hi_w), pack_desc64(b_lo + 2, hi_p), p.idesc, 1u);
              tc_mma<KIND>(d_tmem, pack_desc64(a_lo + 4, hi_w), pack_desc64(b_lo + 4, hi_p), p.idesc, 1u);
              tc_mma<KIND>(d_tmem, pack_desc64(a_lo + 6, hi_w), pack_desc64(b_lo + 6, hi_p), p.idesc, 1u);
            }
            tc_commit(&wempty[ws]);
            if (gidx == p.ngroups - 1) {
              tc_commit(&pempty[ps]);
              if (kc == p.kchunks - 1) tc_commit(&tfull_bar[as]);
            }
          }
          __syncwarp();
          if (++ws == p.n_wbuf) { ws = 0; wph ^= 1; }
        }
        if (++ps == p.n_pbuf) { ps = 0; pph ^= 1; }
      }
      if (++as == 2) { as = 0; aph ^= 1; }
    }
  } else if (warp < HX_WPROD_WARP) {
    // ============================ epilogue (warps 2..9) ============================
    const int q = warp & 3;                        // TMEM lane quarter -> output channels slice*128 + q*32 ..
    const int par = (warp - 2) >> 2;               // 32-column chunks (4 tile rows each) k = par, par + HX_EPI_WARPS/4, ..
    const int L = q * 32 + lane;
    const int blk = L / p.cb, chl = L % p.cb;      // lane block (warp-uniform: cb >= 32) and channel inside the block
    const int oy_g = p.oy[blk], ox_g = p.ox[blk];
    PxStep st;
    st.out_r = p.soy * out.sh; st.out_c = p.sox * out.sw;
    st.add_r = p.soy * add.sh; st.add_c = p.sox * add.sw;
    st.mask_r = p.soy * mask.sh; st.mask_c = p.sox * mask.sw;
    const int jlim = min(p.mj, ox_g < out.w ? (out.w - ox_g + p.sox - 1) / p.sox : 0);
    const int ilim = min(p.mi, oy_g < out.h ? (out.h - oy_g + p.soy - 1) / p.soy : 0);
    int as = 0; unsigned aph = 0;
    double s1 = 0.0, s2 = 0.0;                     // running InstanceNorm sums of (image, channel): flushed when either changes
    int s_img = -1, s_ch = 0;
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      int slice, tj, ti, img;
      hx_tile(p, tile, slice, tj, ti, img);
      const int ch = slice * 128 + chl;
      if (stats && (img != s_img || ch != s_ch)) {
        if (s_img >= 0) { double* srow = stats + ((long long)s_img * p.cstat + s_ch) * 2; atomicAdd(srow, s1); atomicAdd(srow + 1, s2); }
        s1 = s2 = 0.0; s_img = img; s_ch = ch;
      }
      const float b = bias ? bias[ch] : 0.f;
      mbar_wait(&tfull_bar[as], aph);
      tc_fence_after();
      const unsigned taddr0 = tmem_base + ((unsigned)(q * 32) << 16) + (unsigned)(as * 256);
      const int j0 = tj * HX_TW;
      const int nvc = max(0, min(HX_TW, jlim - j0));
#pragma unroll 1
      for (int c0 = par * 32; c0 < 256; c0 += 8 * HX_EPI_WARPS) {
        const int r0 = c0 >> 3;                            // first tile row of this 32-column chunk
        if (r0 >= p.R) break;                              // warp-uniform: short tiles (R < 32) leave columns unused
        float v[32];
        tc_ld32(taddr0 + c0, v);
        const int i0 = ti * p.R + r0;
        const int nvr = max(0, min(min(4, p.R - r0), ilim - i0));
        const int oy = oy_g + p.soy * i0, ox = ox_g + p.sox * j0;
        PxOff off;
        off.out = img * out.sn + oy * out.sh + ox * out.sw;
        off.add = add.ptr ? img * add.sn + oy * add.sh + ox * add.sw : 0;
        off.mask = mask.ptr ? img * mask.sn + oy * mask.sh + ox * mask.sw : 0;
        if (nvr == 4 && nvc == 8)
          px_chunk<8, true>(v, off, st, 4, 8, ch, lane, b, p.flags, add, mask, out, stats != nullptr, s1, s2);
        else if (nvr > 0 && nvc > 0)
          px_chunk<8>(v, off, st, nvr, nvc, ch, lane, b, p.flags, add, mask, out, stats != nullptr, s1, s2);
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[as]);
      if (++as == 2) { as = 0; aph ^= 1; }
    }
    if (stats && s_img >= 0) { double* srow = stats + ((long long)s_img * p.cstat + s_ch) * 2; atomicAdd(srow, s1); atomicAdd(srow + 1, s2); }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// Geometry of one launch in the kernel's general form (plain gather launches and block-stacked ones, include/ast.h).
struct HxGeom {
  int mi, mj, ntaps, flags;
  int nblk, cb, sy, soy, sox, cstat;                         // cstat: channels per image in the statistics array
  int oy[4], ox[4];
  const int16_t* dy;
  const int16_t* dx;
  double* stats;
};

// 1 = launched, 0 = not applicable, other = error.  cout = rows of the packed filter per tap (128 * slices).
static int hx_launch(const ast_image* in, const void* weights, const float* bias, const ast_image* add, const ast_image* mask,
                     const ast_image* out, const HxGeom& g, int cout, cudaStream_t stream) {
  if (!img32_ok(out) || !img32_ok(add) || !img32_ok(mask)) return 0;
  const int esz = in->dtype == AST_F32 ? 4 : 2;
  if ((in->c * esz) % 128 != 0 || cout % 128 != 0) return 0;
  int dy_min = 1 << 30, dy_max = -(1 << 30), dx_min = 1 << 30, dx_max = -(1 << 30);
  for (int t = 0; t < g.ntaps; ++t) {
    dy_min = g.dy[t] < dy_min ? g.dy[t] : dy_min; dy_max = g.dy[t] > dy_max ? g.dy[t] : dy_max;
    dx_min = g.dx[t] < dx_min ? g.dx[t] : dx_min; dx_max = g.dx[t] > dx_max ? g.dx[t] : dx_max;
  }
  if (dx_max - dx_min > 4 || dy_max - dy_min > 8) return 0;
  HxParams p;
  memset(&p, 0, sizeof(p));
  p.kc = 128 / esz;
  p.kchunks = in->c / p.kc;
  if (g.ntaps * p.kchunks < 8) return 0;           // short K loops are epilogue bound: conv_px / conv_tc take them
  EncodeTiledFn encode = get_encode();
  if (!encode) return 0;
  p.mi = g.mi; p.mj = g.mj;
  p.nblk = g.nblk; p.cb = g.cb; p.sy = g.sy; p.soy = g.soy; p.sox = g.sox;
  for (int b = 0; b < 4; ++b) { p.oy[b] = g.oy[b]; p.ox[b] = g.ox[b]; }
  p.cstat = g.nblk == 1 ? cout : g.cb;
  p.ntaps = g.ntaps; p.flags = g.flags; p.cout = cout; p.n_img = in->n; p.n_slices = cout / 128;
  p.dy_min = dy_min; p.dx_min = dx_min;
  for (int t = 0; t < g.ntaps; ++t) { p.tdy[t] = g.dy[t] - dy_min; p.tdx[t] = g.dx[t] - dx_min; }
  p.pw = HX_TW + (dx_max - dx_min);
  const int budget = 225 * 1024 - 1024;
  p.tg = g.ntaps % 3 == 0 ? 3 : (g.ntaps % 2 == 0 ? 2 : (g.ntaps >= 3 ? 3 : g.ntaps));
  p.ngroups = (g.ntaps + p.tg - 1) / p.tg;
  const int w_slot = p.tg * 128 * 128;
  // tile rows: as few row tiles as possible, split evenly (66 rows -> 3 x 22 instead of 32 + 32 + 2); fewer rows per tile
  // when two patches (sy > 1: taller halos) and two weight slots would not fit.  (N = 128 tiles, 5 x 16 rows, were
  // measured SLOWER for the 66-row gradients, 68 vs 50 us: an N = 128 MMA needs 126 B/clk of operand fetch for its 65
  // clk and loses more to the concurrent TMA / epilogue shared-memory traffic than an N = 176 one.)
  auto patch_rows = [&](int R) { return (R - 1) * p.sy + (dy_max - dy_min) + 1; };
  int rmax = 32;
  while (rmax > 2 && (2 * ((p.pw * patch_rows(rmax) * 128 + 1023) & ~1023) + 2 * w_slot > budget || patch_rows(rmax) > 256)) rmax -= 2;
  p.tiles_i = (p.mi + rmax - 1) / rmax;
  p.R = (p.mi + p.tiles_i - 1) / p.tiles_i;
  p.R = (p.R + 1) & ~1;                             // N = 8 R must be a multiple of 16
  p.tiles_i = (p.mi + p.R - 1) / p.R;
  p.tiles_j = (p.mj + HX_TW - 1) / HX_TW;
  p.total_tiles = (long long)p.n_img * p.tiles_i * p.tiles_j * p.n_slices;
  p.ph = patch_rows(p.R);
  p.patch_tx = p.pw * p.ph * 128;
  p.patch_bytes = (p.patch_tx + 1023) & ~1023;
  p.n_pbuf = 2;
  p.n_wbuf = (budget - p.n_pbuf * p.patch_bytes) / w_slot;
  if (p.n_wbuf < 2) return 0;
  if (p.n_wbuf > HX_MAX_WBUF) p.n_wbuf = HX_MAX_WBUF;
  if (budget - p.n_pbuf * p.patch_bytes - p.n_wbuf * w_slot >= p.patch_bytes) p.n_pbuf = 3;
  const unsigned fmt = tc_operand_fmt(in->dtype);
  p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((unsigned)((8 * p.R) >> 3) << 17) | ((128u >> 4) << 24);

  alignas(64) CUtensorMap tm_in, tm_w;
  const CUtensorMapDataType dt = tc_tmap_dtype(in->dtype);
  {
    cuuint64_t dims[4] = {(cuuint64_t)in->c, (cuuint64_t)in->w, (cuuint64_t)in->h, (cuuint64_t)in->n};
    cuuint64_t strides[3] = {(cuuint64_t)in->sw * esz, (cuuint64_t)in->sh * esz, (cuuint64_t)in->sn * esz};
    cuuint32_t box[4] = {(cuuint32_t)p.kc, (cuuint32_t)p.pw, (cuuint32_t)p.ph, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    if (int r = cached_tensor_map(encode, &tm_in, dt, 4, in->ptr, dims, strides, box, estr, CU_TENSOR_MAP_SWIZZLE_128B,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B)) return r;
  }
  {   // packed weights [tap][cout][cin] as a 3-D tensor: a box is tg taps x 128 output channels x one cin-chunk
    cuuint64_t dims[3] = {(cuuint64_t)in->c, (cuuint64_t)cout, (cuuint64_t)g.ntaps};
    cuuint64_t strides[2] = {(cuuint64_t)in->c * esz, (cuuint64_t)cout * in->c * esz};
    cuuint32_t box[3] = {(cuuint32_t)p.kc, 128, (cuuint32_t)p.tg};
    cuuint32_t estr[3] = {1, 1, 1};
    if (int r = cached_tensor_map(encode, &tm_w, dt, 3, const_cast<void*>(weights), dims, strides, box, estr,
                                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B)) return r;
  }
  const size_t smem = 1024 + (size_t)p.n_pbuf * p.patch_bytes + (size_t)p.n_wbuf * w_slot;
  const int grid = (int)(p.total_tiles < num_sms() ? p.total_tiles : num_sms());
  cudaError_t e;
  if (in->dtype != AST_F32) {          // kind::f16 (bf16 or fp16 operands, the format is in the instruction descriptor)
    e = set_max_smem(conv_hx_kernel<0>, smem);
    if (e == cudaSuccess) launch_k(conv_hx_kernel<0>, grid, HX_THREADS, smem, stream, tm_in, tm_w, p, bias, to_img32(add), to_img32(mask), to_img32(out), g.stats);
  } else {
    e = set_max_smem(conv_hx_kernel<1>, smem);
    if (e == cudaSuccess) launch_k(conv_hx_kernel<1>, grid, HX_THREADS, smem, stream, tm_in, tm_w, p, bias, to_img32(add), to_img32(mask), to_img32(out), g.stats);
  }
  if (e != cudaSuccess) { set_error("conv_hx: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); return (int)e; }
  count_launch();
  return 1;
}

// 1 = launched, 0 = not applicable (the caller continues with conv_px / conv_tc), other = error.
int conv_gather_hx(const ast_image* in, const void* weights, const float* bias, const ast_image* add,
                   const ast_image* mask, const ast_image* out, const ast_gather_geom* g, int cpad, bool thin,
                   cudaStream_t stream) {
  static const int enabled = [] { const char* e = getenv("AST_CONV_HX"); return e ? atoi(e) : 1; }();   // A/B switch
  if (!enabled) return 0;
  if (thin || g->pooled || g->w_img_stride != 0 || g->si != 1 || cpad % 128 != 0 || out->c != cpad) return 0;
  int dy_min = 1 << 30, dy_max = -(1 << 30);
  for (int t = 0; t < g->ntaps; ++t) { dy_min = g->dy[t] < dy_min ? g->dy[t] : dy_min; dy_max = g->dy[t] > dy_max ? g->dy[t] : dy_max; }
  if (dy_max - dy_min > 4) return 0;
  HxGeom hg;
  memset(&hg, 0, sizeof(hg));
  hg.mi = g->mi; hg.mj = g->mj; hg.ntaps = g->ntaps; hg.flags = g->flags;
  hg.nblk = 1; hg.cb = 128; hg.sy = 1; hg.soy = hg.sox = g->so;
  hg.oy[0] = g->oy0; hg.ox[0] = g->ox0;
  hg.dy = g->dy; hg.dx = g->dx; hg.stats = g->stats;
  const int r = hx_launch(in, weights, bias, add, mask, out, hg, cpad, stream);
  if (r != 1) return r;
  count_work(FAM_CONV_HX, conv_flops(in, out, g), conv_bytes(in, out, g, add, mask));
  AST_CUDA_LAUNCH_CHECK();
  return 1;
}

// Block-stacked launches whose filter does not fit in shared memory (conv_st.cu keeps it resident): the same kernel with
// the 128 lanes split into g->nblk output blocks and the filter streamed.  1 = launched, 0 = not applicable.
int conv_stacked_hx(const ast_image* in, const void* weights, const float* bias, const ast_image* add,
                    const ast_image* mask, const ast_image* out, const ast_stacked_geom* g, cudaStream_t stream) {
  HxGeom hg;
  memset(&hg, 0, sizeof(hg));
  hg.mi = g->mi; hg.mj = g->mj; hg.ntaps = g->nvt; hg.flags = g->flags;
  hg.nblk = g->nblk; hg.cb = 128 / g->nblk; hg.sy = g->sy; hg.soy = g->soy; hg.sox = g->sox;
  for (int b = 0; b < g->nblk; ++b) { hg.oy[b] = g->oy[b]; hg.ox[b] = g->ox[b]; }
  hg.dy = g->dy; hg.dx = g->dx; hg.stats = g->stats;
  return hx_launch(in, weights, bias, add, mask, out, hg, 128, stream);
}

}  // namespace ast
