// Block-stacked, pixels-as-N tcgen05 gather convolution for THIN outputs (32 or 64 output channels): ast_conv_stacked.
//
// The 32- and 64-channel 256^2 layers were the slowest convolutions of the step per FLOP (conv_ws.cu, pixels as M: an
// M128 x N32 MMA over 64-byte rows costs ~160 clk for 1/8 of the tensor core's work).  With pixels as N the output
// channels become the M = 128 TMEM lanes, and a thin layer fills them with nblk = 128/cout independent BLOCKS that read the
// same input pixels:
//   * the sub-pixel phases of a stride-2 ConvTranspose2d / stride-2 data gradient: the 1+2+2+4 taps of the four phases
//     are 4 distinct input shifts -> 4 virtual taps instead of 9 tap MMAs, each with 4x the N;
//   * row-interleaved stride-1 convolutions: block g owns output rows g, g+nblk, ..; virtual tap v = dy + g, so a 9-tap
//     vertical layer costs 12 MMAs per 4 output rows instead of 36.
//
//     D^T[(g, co)][pixel (r, c)] += Ws[v][g*cout + co][k] * Patch[(r*sy + dy_v) * pw + c + dx_v][k]
//
// The whole stacked filter ([nvt][128][cin], zero rows where a block has no tap) stays resident in shared memory; one halo
// patch per tile (and cin-chunk) serves every virtual tap through shifted descriptors whose 8-pixel groups are
// SBO = sy patch rows apart.  Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-9 = epilogue (one thread =
// one (block, channel) lane x 32 pixels of a tcgen05.ld, px_common.cuh).
// Measured at B=32, 256^2 (scratch/prof_st.py, profiles/r02_summary.md; the same layers on conv_ws.cu before):
//   first 9x9 layer (9 vertical taps, +stats) 209 -> 86 us, ConvTranspose 64->32 236 -> 77 us, stride-2 data gradient
//   64->32 256 -> 76 us, last layer forward 175 -> 83 us / data gradient 201 -> 92 us, VGG conv1_1 196 -> 137 us and
//   its data gradient 161 -> 95 us, ConvTranspose 128->64 104 -> 65 us, stride-2 data gradient 128->64 112 -> 55 us.
//   With the epilogue removed the first layer takes 56 us (MMA), with the MMAs removed 62 us (epilogue), with both 30 us
//   (TMA): the kernel runs at the overlap of an MMA-bound and an epilogue-bound pipeline, ~2x its HBM floor.
#include "px_common.cuh"

namespace ast {

constexpr int ST_EPI_WARPS = 8;    // 2 per TMEM lane quarter (16 were measured slower: 96 registers with spills, first layer 129 -> 226 us)
constexpr int ST_THREADS = 64 + 32 * ST_EPI_WARPS;
constexpr int ST_MAX_PBUF = 4;
constexpr int ST_TW = 8;

struct StParams {
  int mi, mj, tiles_i, tiles_j, n_img, R;
  int nblk, cb, nvt, kchunks, kc, rowb, flags;
  int sy, soy, sox;
  int oy[4], ox[4];
  int dy_min, dx_min, ph, pw;
  int patch_bytes, patch_tx, n_pbuf, w_tile_bytes, w_total_bytes;
  unsigned idesc, layout_type;
  long long total_tiles;
  short tdy[AST_MAX_VTAPS];
  short tdx[AST_MAX_VTAPS];
};

__device__ __forceinline__ void st_tile(const StParams& p, long long tile, int& tj, int& ti, int& img) {
  long long r = tile;
  tj = (int)(r % p.tiles_j); r /= p.tiles_j;
  ti = (int)(r % p.tiles_i);
  img = (int)(r / p.tiles_i);
}

template <int KIND>
__global__ void __launch_bounds__(ST_THREADS, 1)
conv_st_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_w, const StParams p,
               const float* __restrict__ bias, const Img32 add, const Img32 mask, const Img32 out,
               double* __restrict__ stats) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long pfull[ST_MAX_PBUF], pempty[ST_MAX_PBUF], tfull_bar[2], tempty_bar[2], wbar;
  __shared__ unsigned tmem_slot;
  __shared__ unsigned s_tapoff[AST_MAX_VTAPS];   // per virtual tap: start offset of the pixel operand inside the patch, 16-byte units

  unsigned char* smem_w = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char* smem_p = smem_w + ((p.w_total_bytes + 1023) & ~1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_in) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_w) : "memory");
    for (int s = 0; s < p.n_pbuf; ++s) { mbar_init(&pfull[s], 1); mbar_init(&pempty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 32 * ST_EPI_WARPS); }
    mbar_init(&wbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (threadIdx.x >= 64 && threadIdx.x - 64 < p.nvt) {
    const int t = threadIdx.x - 64;
    s_tapoff[t] = (unsigned)((p.tdy[t] * p.pw + p.tdx[t]) * p.rowb) >> 4;
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
    // ============================ TMA producer: the stacked filter once, then one patch per (tile, cin-chunk) ============================
    if (lane == 0) {
      mbar_expect_tx(&wbar, (unsigned)p.w_total_bytes);
      for (int v = 0; v < p.nvt; ++v)
        for (int kc = 0; kc < p.kchunks; ++kc)
          tma_load_2d(smem_w + (size_t)(v * p.kchunks + kc) * p.w_tile_bytes, &tm_w, &wbar, kc * p.kc, v * 128);
      int s = 0; unsigned ph = 0;
      for (long long tile = tile_begin(p.total_tiles), tile_e = tile_end(p.total_tiles); tile < tile_e; ++tile) {
        int tj, ti, img;
        st_tile(p, tile, tj, ti, img);
        const int x0 = tj * ST_TW + p.dx_min, y0 = ti * p.R * p.sy + p.dy_min;
        for (int kc = 0; kc < p.kchunks; ++kc) {
          mbar_wait(&pempty[s], ph ^ 1);
          mbar_expect_tx(&pfull[s], (unsigned)p.patch_tx);
          tma_load_4d(smem_p + (size_t)s * p.patch_bytes, &tm_in, &pfull[s], kc * p.kc, x0, y0, img);
          if (++s == p.n_pbuf) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    int s = 0; unsigned ph = 0; int as = 0; unsigned aph = 0;
    const int kmma = p.rowb / 32;
    mbar_wait(&wbar, 0);
    tc_fence_after();
    const unsigned hi_w = ((8u * (unsigned)p.rowb) >> 4) | (1u << 14) | (p.layout_type << 29);                     // filter: dense rows
    const unsigned hi_p = (((unsigned)(p.sy * p.pw * p.rowb)) >> 4) | (1u << 14) | (p.layout_type << 29);          // patch: 8-pixel groups sy rows apart
    const unsigned w_lo0 = ((smem_u32(smem_w) & 0x3FFFFu) >> 4) | (1u << 16);
    const unsigned w16 = (unsigned)p.w_tile_bytes >> 4;
    for (long long tile = tile_begin(p.total_tiles), tile_e = tile_end(p.total_tiles); tile < tile_e; ++tile) {
      mbar_wait(&tempty_bar[as], aph ^ 1);
      tc_fence_after();
      const unsigned d_tmem = tmem_base + (unsigned)(as * 256);
      for (int kc = 0; kc < p.kchunks; ++kc) {
        mbar_wait(&pfull[s], ph);
        tc_fence_after();
        if (lane == 0) {
          const unsigned p_lo = ((smem_u32(smem_p + (size_t)s * p.patch_bytes) & 0x3FFFFu) >> 4) | (1u << 16);
          unsigned a_lo = w_lo0 + (unsigned)kc * w16;
          const unsigned a_step = (unsigned)p.kchunks * w16;
          unsigned acc = kc > 0 ? 1u : 0u;
#pragma unroll 2
          for (int v = 0; v < p.nvt; ++v) {
            const unsigned b_lo = p_lo + s_tapoff[v];
            tc_mma<KIND>(d_tmem, pack_desc64(a_lo, hi_w), pack_desc64(b_lo, hi_p), p.idesc, acc);
            tc_mma<KIND>(d_tmem, pack_desc64(a_lo + 2, hi_w), pack_desc64(b_lo + 2, hi_p), p.idesc, 1u);
            if (kmma == 4) {
              tc_mma<KIND>(d_tmem, pack_desc64(a_lo + 4, hi_w), pack_desc64(b_lo + 4, hi_p), p.idesc, 1u);
              tc_mma<KIND>(d_tmem, pack_desc64(a_lo + 6, hi_w), pack_desc64(b_lo + 6, hi_p), p.idesc, 1u);
            }
            acc = 1u;
            a_lo += a_step;
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
    // ============================ epilogue (warps 2..9) ============================
    const int q = warp & 3;                        // TMEM lane quarter
    const int par = (warp - 2) >> 2;               // 32-column chunks (4 grid rows each) k = par, par + ST_EPI_WARPS/4, ..
    const int L = q * 32 + lane;
    const int g = L / p.cb, ch = L % p.cb;         // lane block (warp-uniform: cb >= 32) and output channel
    const int oy_g = p.oy[g], ox_g = p.ox[g];
    PxStep st;
    st.out_r = p.soy * out.sh; st.out_c = p.sox * out.sw;
    st.add_r = p.soy * add.sh; st.add_c = p.sox * add.sw;
    st.mask_r = p.soy * mask.sh; st.mask_c = p.sox * mask.sw;
    // grid rows / columns whose output pixel of THIS block lies inside the image
    const int ilim = min(p.mi, oy_g < out.h ? (out.h - oy_g + p.soy - 1) / p.soy : 0);
    const int jlim = min(p.mj, ox_g < out.w ? (out.w - ox_g + p.sox - 1) / p.sox : 0);
    const float b = bias ? bias[ch] : 0.f;
    int as = 0; unsigned aph = 0;
    double s1 = 0.0, s2 = 0.0;                     // running InstanceNorm sums of (image, channel): flushed when the image changes
    int s_img = -1;
    const int nchunks = (p.R + 3) >> 2;
    for (long long tile = tile_begin(p.total_tiles), tile_e = tile_end(p.total_tiles); tile < tile_e; ++tile) {
      int tj, ti, img;
      st_tile(p, tile, tj, ti, img);
      if (stats && img != s_img) {
        if (s_img >= 0) { double* srow = stats + ((long long)s_img * p.cb + ch) * 2; atomicAdd(srow, s1); atomicAdd(srow + 1, s2); }
        s1 = s2 = 0.0; s_img = img;
      }
      mbar_wait(&tfull_bar[as], aph);
      tc_fence_after();
      const unsigned taddr0 = tmem_base + ((unsigned)(q * 32) << 16) + (unsigned)(as * 256);
      const int j0 = tj * ST_TW;
      const int nvc = max(0, min(ST_TW, jlim - j0));
#pragma unroll 1
      for (int k = par; k < nchunks; k += ST_EPI_WARPS / 4) {
        float v[32];
        tc_ld32(taddr0 + (unsigned)(k * 32), v);
        const int r0 = k * 4;
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
    if (stats && s_img >= 0) { double* srow = stats + ((long long)s_img * p.cb + ch) * 2; atomicAdd(srow, s1); atomicAdd(srow + 1, s2); }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace ast

using namespace ast;

namespace ast {
int conv_stacked_hx(const ast_image* in, const void* weights, const float* bias, const ast_image* add,
                    const ast_image* mask, const ast_image* out, const ast_stacked_geom* g, cudaStream_t stream);   // conv_hx.cu
}

extern "C" int ast_conv_stacked(const ast_image* in, const void* weights, const float* bias, const ast_image* add,
                                const ast_image* mask, const ast_image* out, const ast_stacked_geom* g, void* stream) {
  AST_CHECK_ARG(in && weights && out && g, "ast_conv_stacked: null argument");
  AST_CHECK_ARG(g->nblk == 2 || g->nblk == 4, "ast_conv_stacked: nblk must be 2 or 4 (got %d)", g->nblk);
  AST_CHECK_ARG(g->nvt >= 1 && g->nvt <= AST_MAX_VTAPS, "ast_conv_stacked: nvt %d out of range", g->nvt);
  AST_CHECK_ARG(in->n == out->n, "ast_conv_stacked: batch mismatch %d vs %d", in->n, out->n);
  AST_CHECK_ARG(g->mi > 0 && g->mj > 0 && g->sy >= 1 && g->soy >= 1 && g->sox >= 1, "ast_conv_stacked: bad geometry");
  AST_CHECK_ARG(in->dtype == AST_F32 || in->dtype == AST_BF16 || in->dtype == AST_F16, "ast_conv_stacked: bad input dtype");
  AST_CHECK_ARG(out->dtype == AST_F32 || out->dtype == AST_BF16 || out->dtype == AST_F16, "ast_conv_stacked: bad output dtype");
  AST_CHECK_ARG(!add || same_shape(add, out), "ast_conv_stacked: add image shape mismatch");
  AST_CHECK_ARG(!mask || same_shape(mask, out), "ast_conv_stacked: mask image shape mismatch");
  const int cb = 128 / g->nblk;
  AST_CHECK_ARG(out->c == cb && out->sc == 1, "ast_conv_stacked: out must have %d contiguous channels (got %d)", cb, out->c);
  AST_CHECK_ARG(img32_ok(out) && img32_ok(add) && img32_ok(mask), "ast_conv_stacked: output images must span < 2^31 elements, sc = 1");
  AST_CHECK_ARG(in->sc == 1, "ast_conv_stacked: input must be channel-contiguous");
  if (in->n == 0) return 0;
  const int esz = in->dtype == AST_F32 ? 4 : 2;
  const int cbytes = in->c * esz;
  AST_CHECK_ARG(cbytes == 64 || cbytes % 128 == 0, "ast_conv_stacked: input pixels must be 64 bytes or a multiple of 128 bytes (got %d)", cbytes);
  int dy_min = 1 << 30, dy_max = -(1 << 30), dx_min = 1 << 30, dx_max = -(1 << 30);
  for (int t = 0; t < g->nvt; ++t) {
    dy_min = g->dy[t] < dy_min ? g->dy[t] : dy_min; dy_max = g->dy[t] > dy_max ? g->dy[t] : dy_max;
    dx_min = g->dx[t] < dx_min ? g->dx[t] : dx_min; dx_max = g->dx[t] > dx_max ? g->dx[t] : dx_max;
  }
  AST_CHECK_ARG(dx_max - dx_min <= 8, "ast_conv_stacked: horizontal tap span %d too wide", dx_max - dx_min);
  StParams p;
  memset(&p, 0, sizeof(p));
  p.rowb = cbytes == 64 ? 64 : 128;
  p.kc = p.rowb / esz;
  p.kchunks = in->c / p.kc;
  p.nblk = g->nblk; p.cb = cb; p.nvt = g->nvt; p.flags = g->flags;
  p.mi = g->mi; p.mj = g->mj; p.sy = g->sy; p.soy = g->soy; p.sox = g->sox; p.n_img = in->n;
  for (int b = 0; b < 4; ++b) { p.oy[b] = b < g->nblk ? g->oy[b] : 0; p.ox[b] = b < g->nblk ? g->ox[b] : 0; }
  p.dy_min = dy_min; p.dx_min = dx_min;
  for (int t = 0; t < g->nvt; ++t) { p.tdy[t] = g->dy[t] - dy_min; p.tdx[t] = g->dx[t] - dx_min; }
  p.w_tile_bytes = 128 * p.rowb;
  p.w_total_bytes = g->nvt * p.kchunks * p.w_tile_bytes;
  p.pw = ST_TW + (dx_max - dx_min);
  const int budget = 225 * 1024 - 1024 - ((p.w_total_bytes + 1023) & ~1023);
  // grid rows per tile (N = 8R): the largest even R <= 32 that leaves room for three patch buffers (two when even R = 8
  // does not), then split mi evenly (66 rows -> 3 x 22, not 32 + 32 + 2)
  auto patch_bytes = [&](int R) { return (p.pw * ((R - 1) * p.sy + (dy_max - dy_min) + 1) * p.rowb + 1023) & ~1023; };
  int rmax = 0;
  for (int nb = 3; nb >= 2 && !rmax; --nb)
    for (int R = 32; R >= (nb == 3 ? 16 : 2); R -= 2)
      if (nb * patch_bytes(R) <= budget && (R - 1) * p.sy + (dy_max - dy_min) + 1 <= 256) { rmax = R; break; }
  const double taps = (double)g->ntaps;          // real filter taps over all blocks (zero rows of the stacked filter do not count)
  const double pix = (double)in->n * g->mi * g->mj;
  const double flops = 2.0 * pix * taps * in->c * cb;
  double pix_in = (double)g->mi * g->sy * g->mj;
  if (pix_in > (double)in->h * in->w) pix_in = (double)in->h * in->w;
  const double bytes = in->n * pix_in * in->c * esize(in) +
                       pix * g->nblk * cb * (esize(out) + (add ? esize(add) : 0) + (mask ? esize(mask) : 0));
  if (rmax < 12) {         // the filter leaves no room for two useful patches: stream it (conv_hx.cu)
    const int r = conv_stacked_hx(in, weights, bias, add, mask, out, g, (cudaStream_t)stream);
    AST_CHECK_ARG(r != 0 || rmax > 0, "ast_conv_stacked: the stacked filter (%d bytes) neither fits in shared memory nor suits the streaming kernel", p.w_total_bytes);
    if (r == 1) {
      count_work(FAM_CONV_HX, flops, bytes);
      AST_CUDA_LAUNCH_CHECK();
      return 0;
    }
    if (r != 0) return r;
  }
  p.tiles_i = (p.mi + rmax - 1) / rmax;
  p.R = (p.mi + p.tiles_i - 1) / p.tiles_i;
  p.R = (p.R + 1) & ~1;                              // N = 8R must be a multiple of 16
  p.tiles_i = (p.mi + p.R - 1) / p.R;
  p.tiles_j = (p.mj + ST_TW - 1) / ST_TW;
  p.total_tiles = (long long)p.n_img * p.tiles_i * p.tiles_j;
  p.ph = (p.R - 1) * p.sy + (dy_max - dy_min) + 1;
  p.patch_tx = p.pw * p.ph * p.rowb;
  p.patch_bytes = (p.patch_tx + 1023) & ~1023;
  p.n_pbuf = budget / p.patch_bytes;
  if (p.n_pbuf > ST_MAX_PBUF) p.n_pbuf = ST_MAX_PBUF;
  p.layout_type = p.rowb == 128 ? 2u : 4u;
  const unsigned fmt = tc_operand_fmt(in->dtype);
  p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((unsigned)((8 * p.R) >> 3) << 17) | ((128u >> 4) << 24);

  EncodeTiledFn encode = get_encode();
  AST_CHECK_ARG(encode, "ast_conv_stacked: cuTensorMapEncodeTiled entry point not available");
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
    cuuint64_t dims[2] = {(cuuint64_t)in->c, (cuuint64_t)((long long)g->nvt * 128)};
    cuuint64_t strides[1] = {(cuuint64_t)in->c * esz};
    cuuint32_t box[2] = {(cuuint32_t)p.kc, 128};
    cuuint32_t estr[2] = {1, 1};
    if (int r = cached_tensor_map(encode, &tm_w, dt, 2, const_cast<void*>(weights), dims, strides, box, estr, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B)) return r;
  }
  const size_t smem = 1024 + ((p.w_total_bytes + 1023) & ~1023) + (size_t)p.n_pbuf * p.patch_bytes;
  const int grid = (int)(p.total_tiles < num_sms() ? p.total_tiles : num_sms());
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e;
  if (in->dtype != AST_F32) {          // kind::f16 (bf16 or fp16 operands, the format is in the instruction descriptor)
    e = set_max_smem(conv_st_kernel<0>, smem);
    if (e == cudaSuccess) launch_k(conv_st_kernel<0>, grid, ST_THREADS, smem, s, tm_in, tm_w, p, bias, to_img32(add), to_img32(mask), to_img32(out), g->stats);
  } else {
    e = set_max_smem(conv_st_kernel<1>, smem);
    if (e == cudaSuccess) launch_k(conv_st_kernel<1>, grid, ST_THREADS, smem, s, tm_in, tm_w, p, bias, to_img32(add), to_img32(mask), to_img32(out), g->stats);
  }
  if (e != cudaSuccess) { set_error("ast_conv_stacked: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); return (int)e; }
  count_launch();
  count_work(FAM_CONV_ST, flops, bytes);
  AST_CUDA_LAUNCH_CHECK();
  return 0;
}
