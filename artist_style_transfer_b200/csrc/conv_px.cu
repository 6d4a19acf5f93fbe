// "Pixels-as-N" orientation of the tcgen05 gather convolution, for layers with at most 128 output channels.
//
// Measured on B200 (scratch/mma_bench.cu, profiles/r01_summary.md): a tcgen05.mma of cta_group::1 costs
// max(~64, N/2) cycles, and ~104 cycles as soon as tcgen05.commit is used to recycle pipeline stages - whatever M is.
// conv_tc.cu maps output channels to N, so a 128-channel layer pays 104 cycles for a 128x128x(32 B) MMA and a
// 64-channel layer the same for half the work.  Here the operands swap roles:
//
//     D^T[co][pixel] (fp32, TMEM: lane = output channel, column = pixel)  +=  W_t[co][k] * X[pixel + tap_t][k]
//
//   A operand = the packed weights of tap t (<= 128 rows, K-major),  B operand = 256 pixels (two 128-pixel tiles of the
//   same image, K-major rows of the NHWC tensor), so every MMA is 128 x 256 x (32 B) and runs at the tensor peak.
//
// Warp roles as in conv_tc.cu (warp 0 TMA, warp 1 MMA, warps 2-9 epilogue).  In the epilogue a thread owns ONE output
// channel and 32 consecutive pixels of a tcgen05.ld: the bias is a scalar, the InstanceNorm statistics (sum x, sum x^2)
// are plain per-thread sums, and for a fixed pixel the 32 lanes of a warp touch 32 consecutive channels, so the NHWC
// stores and the add / mask loads are coalesced without a shared-memory transpose.
#include "px_common.cuh"

namespace ast {

constexpr int PX_THREADS = 320;
constexpr int PX_MAX_STAGES = 8;
constexpr int PX_PRODUCERS = 4;    // producer lanes (one thread issues only ~1 TMA instruction per 150-200 cycles)

struct PxParams {
  int mi, mj, tw, th, tiles_i, tiles_j, n_img;
  int ntaps, kchunks, kc;
  int si, so, oy0, ox0;
  int cout, flags;                 // cout = weight rows per tap (64 or 128)
  int w_rows_per_img;
  int stages, w_bytes, a_bytes, stage_bytes, rowb;
  int pairs_per_img;
  unsigned idesc, layout_type, sbo;
  long long total_pairs;
  short dy[AST_MAX_TAPS];
  short dx[AST_MAX_TAPS];
};

template <int KIND>
__global__ void __launch_bounds__(PX_THREADS, 1)
conv_px_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_w, const PxParams p,
               const float* __restrict__ bias, const Img32 add, const Img32 mask, const Img32 out,
               double* __restrict__ stats) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long full_bar[PX_MAX_STAGES], empty_bar[PX_MAX_STAGES], tfull_bar[2], tempty_bar[2];
  __shared__ unsigned tmem_slot;
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);      // the stage ring
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ksteps = p.ntaps * p.kchunks;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_in) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_w) : "memory");
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 256); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
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
    // ============================ TMA producers (lanes 0..PX_PRODUCERS-1 take the stages round-robin) ============
    // a lane must never run two ring cycles ahead of the consumer (the 1-bit mbarrier parity would alias): at most
    // `stages` lanes take part, then consecutive items of one lane are <= one cycle apart
    const int nprod = p.stages < PX_PRODUCERS ? p.stages : PX_PRODUCERS;
    if (lane < nprod) {
      int s = 0, turn = 0; unsigned ph = 0;
      for (long long pair = blockIdx.x; pair < p.total_pairs; pair += gridDim.x) {
        const int img = (int)(pair / p.pairs_per_img);
        const int sp0 = 2 * (int)(pair % p.pairs_per_img);
        // two consecutive spatial tiles (row-major over ti, tj); a tile index past the end gives ti == tiles_i: the
        // loads are fully out of bounds (zero fill) and the epilogue stores nothing
        const int ti0 = sp0 / p.tiles_j, tj0 = sp0 - ti0 * p.tiles_j;
        const int ti1 = (sp0 + 1) / p.tiles_j, tj1 = (sp0 + 1) - ti1 * p.tiles_j;
        const int x0 = p.si * tj0 * p.tw, y0 = p.si * ti0 * p.th, x1 = p.si * tj1 * p.tw, y1 = p.si * ti1 * p.th;
        const int wrow0 = img * p.w_rows_per_img;
        for (int t = 0; t < p.ntaps; ++t) {
          for (int kc = 0; kc < p.kchunks; ++kc) {
            if (turn == lane) {
              mbar_wait(&empty_bar[s], ph ^ 1);
              unsigned char* sw = smem + (size_t)s * p.stage_bytes;
              mbar_expect_tx(&full_bar[s], (unsigned)p.stage_bytes);
              tma_load_2d(sw, &tm_w, &full_bar[s], kc * p.kc, wrow0 + t * p.cout);
              tma_load_4d(sw + p.w_bytes, &tm_in, &full_bar[s], kc * p.kc, x0 + p.dx[t], y0 + p.dy[t], img);
              tma_load_4d(sw + p.w_bytes + p.a_bytes, &tm_in, &full_bar[s], kc * p.kc, x1 + p.dx[t], y1 + p.dy[t], img);
            }
            if (++turn == nprod) turn = 0;
            if (++s == p.stages) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    int s = 0; unsigned ph = 0; int as = 0; unsigned aph = 0;
    const int kmma = p.rowb / 32;
    const unsigned desc_hi = (p.sbo >> 4) | (1u << 14) | (p.layout_type << 29);
    for (long long pair = blockIdx.x; pair < p.total_pairs; pair += gridDim.x) {
      mbar_wait(&tempty_bar[as], aph ^ 1);
      tc_fence_after();
      const unsigned d_tmem = tmem_base + (unsigned)(as * 256);
      for (int ks = 0; ks < ksteps; ++ks) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        if (lane == 0) {
          const unsigned w_addr = smem_u32(smem + (size_t)s * p.stage_bytes);
          const unsigned a_lo = ((w_addr & 0x3FFFFu) >> 4) | (1u << 16);                     // A operand: this tap's weights
          const unsigned b_lo = (((w_addr + p.w_bytes) & 0x3FFFFu) >> 4) | (1u << 16);      // B operand: 256 pixels
          tc_mma<KIND>(d_tmem, pack_desc64(a_lo, desc_hi), pack_desc64(b_lo, desc_hi), p.idesc, ks > 0 ? 1u : 0u);
          tc_mma<KIND>(d_tmem, pack_desc64(a_lo + 2, desc_hi), pack_desc64(b_lo + 2, desc_hi), p.idesc, 1u);
          if (kmma == 4) {
            tc_mma<KIND>(d_tmem, pack_desc64(a_lo + 4, desc_hi), pack_desc64(b_lo + 4, desc_hi), p.idesc, 1u);
            tc_mma<KIND>(d_tmem, pack_desc64(a_lo + 6, desc_hi), pack_desc64(b_lo + 6, desc_hi), p.idesc, 1u);
          }
          tc_commit(&empty_bar[s]);
          if (ks == ksteps - 1) tc_commit(&tfull_bar[as]);
        }
        __syncwarp();
        if (++s == p.stages) { s = 0; ph ^= 1; }
      }
      if (++as == 2) { as = 0; aph ^= 1; }
    }
  } else {
    // ============================ epilogue (warps 2..9) ============================
    const int q = warp & 3;                        // TMEM lane quarter -> output channels q*32 .. q*32+31
    const int half = (warp - 2) >> 2;              // which 128-pixel tile of the pair (TMEM columns half*128 ..)
    const int ch = q * 32 + lane;
    const bool ch_ok = ch < p.cout;
    const float b = (bias && ch_ok) ? bias[ch] : 0.f;
    int as = 0; unsigned aph = 0;
    double s1 = 0.0, s2 = 0.0;                     // running InstanceNorm sums of this thread's channel over one image
    int s_img = -1;
    for (long long pair = blockIdx.x; pair < p.total_pairs; pair += gridDim.x) {
      const int img = (int)(pair / p.pairs_per_img);
      if (stats && img != s_img) {                 // flush once per image, not once per tile (atomic contention)
        if (s_img >= 0 && ch_ok) { double* srow = stats + ((long long)s_img * p.cout + ch) * 2; atomicAdd(srow, s1); atomicAdd(srow + 1, s2); }
        s1 = s2 = 0.0; s_img = img;
      }
      const int sp = 2 * (int)(pair % p.pairs_per_img) + half;
      const int ti = sp / p.tiles_j, tj = sp - ti * p.tiles_j;
      mbar_wait(&tfull_bar[as], aph);
      tc_fence_after();
      const unsigned taddr0 = tmem_base + ((unsigned)(q * 32) << 16) + (unsigned)(as * 256 + half * 128);
      if (q * 32 < p.cout) {                       // warp-uniform: quarters beyond cout hold nothing
        const bool fast = (p.tw & 31) == 0;
        for (int c0 = 0; c0 < 128; c0 += 32) {
          float v[32];
          tc_ld32(taddr0 + c0, v);
          if (fast) {           // the chunk is 32 consecutive x positions of tile row c0 / tw
            const int ty = c0 / p.tw, tx0 = c0 - ty * p.tw;
            const int i = ti * p.th + ty, j0 = tj * p.tw + tx0;
            const int oy = p.oy0 + p.so * i, ox = p.ox0 + p.so * j0;
            int nvalid = 0;
            if (i < p.mi && oy < out.h) {
              const int jlim = min(p.mj, (out.w - p.ox0 + p.so - 1) / p.so);    // j < jlim  <=>  j < mj and ox < out.w
              nvalid = max(0, min(32, jlim - j0));
            }
            PxOff off;
            off.out = img * out.sn + oy * out.sh + ox * out.sw;
            off.add = add.ptr ? img * add.sn + oy * add.sh + ox * add.sw : 0;
            off.mask = mask.ptr ? img * mask.sn + oy * mask.sh + ox * mask.sw : 0;
            PxStep st;
            st.out_r = st.add_r = st.mask_r = 0;
            st.out_c = p.so * out.sw; st.add_c = p.so * add.sw; st.mask_c = p.so * mask.sw;
            if (nvalid > 0) px_chunk<32>(v, off, st, 1, nvalid, ch, lane, b, p.flags, add, mask, out, stats != nullptr, s1, s2);
          } else {              // lane L describes pixel c0 + L of the tile
            const int row = c0 + lane;
            const int ty = row / p.tw, tx = row - ty * p.tw;
            const int i = ti * p.th + ty, j = tj * p.tw + tx;
            const int oy = p.oy0 + p.so * i, ox = p.ox0 + p.so * j;
            const bool valid = i < p.mi && j < p.mj && oy < out.h && ox < out.w;
            PxOff off;
            off.out = valid ? img * out.sn + oy * out.sh + ox * out.sw : -1;
            off.add = add.ptr ? img * add.sn + oy * add.sh + ox * add.sw : 0;
            off.mask = mask.ptr ? img * mask.sn + oy * mask.sh + ox * mask.sw : 0;
            PxStep st = {0, 0, 0, 0, 0, 0};
            px_chunk<0>(v, off, st, 0, 0, ch, lane, b, p.flags, add, mask, out, stats != nullptr, s1, s2);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[as]);
      if (++as == 2) { as = 0; aph ^= 1; }
    }
    if (stats && s_img >= 0 && ch_ok) { double* srow = stats + ((long long)s_img * p.cout + ch) * 2; atomicAdd(srow, s1); atomicAdd(srow + 1, s2); }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// 1 = launched, 0 = not applicable (the caller falls back to conv_tc), other = error.
// The caller (conv_gather_tc) has validated pointers / alignment and offered the launch to conv_ws first.
// (A resident-weight form was measured slower for every layer it applied to and removed:
//  scratch/dead_variants/conv_px_r01.cu.txt, profiles/r01_summary.md finding 11.)
int conv_gather_px(const ast_image* in, const void* weights, const float* bias, const ast_image* add,
                   const ast_image* mask, const ast_image* out, const ast_gather_geom* g, int cpad, bool thin,
                   cudaStream_t stream) {
  if (thin || g->pooled || (cpad != 64 && cpad != 128) || out->c != cpad) return 0;
  if (!img32_ok(out) || !img32_ok(add) || !img32_ok(mask)) return 0;
  if (in->dtype == AST_F16) return 0;
  const int esz = in->dtype == AST_F32 ? 4 : 2;
  EncodeTiledFn encode = get_encode();
  if (!encode) return 0;

  PxParams p;
  memset(&p, 0, sizeof(p));
  p.mi = g->mi; p.mj = g->mj; p.si = g->si; p.so = g->so; p.oy0 = g->oy0; p.ox0 = g->ox0;
  p.ntaps = g->ntaps; p.flags = g->flags; p.cout = cpad; p.n_img = in->n;
  for (int t = 0; t < g->ntaps; ++t) { p.dy[t] = g->dy[t]; p.dx[t] = g->dx[t]; }
  pick_tile(p.mi, p.mj, 128, &p.tw, &p.th);
  p.tiles_i = (p.mi + p.th - 1) / p.th;
  p.tiles_j = (p.mj + p.tw - 1) / p.tw;
  p.rowb = (in->c * esz) % 128 == 0 ? 128 : 64;
  p.kc = p.rowb / esz;
  p.kchunks = in->c / p.kc;
  if (p.ntaps * p.kchunks * (p.rowb / 32) < 16) return 0;   // short K loops are epilogue bound: conv_tc's vector stores win
  p.w_rows_per_img = 0;
  if (g->w_img_stride) {
    if (g->w_img_stride != (int64_t)g->ntaps * cpad * in->c) return 0;
    p.w_rows_per_img = g->ntaps * cpad;
  }
  p.w_bytes = 128 * p.rowb;          // always a 128-row box: rows >= cout belong to the next tap (or are OOB zeros) and only
  p.a_bytes = 128 * p.rowb;          // feed accumulator lanes that the epilogue ignores
  p.stage_bytes = p.w_bytes + 2 * p.a_bytes;
  p.stages = (200 * 1024) / p.stage_bytes;
  if (p.stages > PX_MAX_STAGES) p.stages = PX_MAX_STAGES;
  p.layout_type = p.rowb == 128 ? 2u : 4u;
  p.sbo = 8u * p.rowb;
  const unsigned fmt = tc_operand_fmt(in->dtype);
  p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
  p.pairs_per_img = (p.tiles_i * p.tiles_j + 1) / 2;
  p.total_pairs = (long long)p.n_img * p.pairs_per_img;

  alignas(64) CUtensorMap tm_in, tm_w;
  const CUtensorMapDataType dt = tc_tmap_dtype(in->dtype);
  const CUtensorMapSwizzle sw = p.rowb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  {
    cuuint64_t dims[4] = {(cuuint64_t)in->c, (cuuint64_t)in->w, (cuuint64_t)in->h, (cuuint64_t)in->n};
    cuuint64_t strides[3] = {(cuuint64_t)in->sw * esz, (cuuint64_t)in->sh * esz, (cuuint64_t)in->sn * esz};
    cuuint32_t box[4] = {(cuuint32_t)p.kc, (cuuint32_t)(p.tw * p.si), (cuuint32_t)(p.th * p.si), 1};
    cuuint32_t estr[4] = {1, (cuuint32_t)p.si, (cuuint32_t)p.si, 1};
    if (int r = cached_tensor_map(encode, &tm_in, dt, 4, in->ptr, dims, strides, box, estr, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B)) return r;
  }
  {
    const long long rows = (long long)g->ntaps * cpad * (g->w_img_stride ? in->n : 1);
    cuuint64_t dims[2] = {(cuuint64_t)in->c, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)in->c * esz};
    cuuint32_t box[2] = {(cuuint32_t)p.kc, 128};
    cuuint32_t estr[2] = {1, 1};
    if (int r = cached_tensor_map(encode, &tm_w, dt, 2, const_cast<void*>(weights), dims, strides, box, estr, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B)) return r;
  }
  const size_t smem = (size_t)p.stages * p.stage_bytes + 1024;
  const int grid = (int)(p.total_pairs < num_sms() ? p.total_pairs : num_sms());
  cudaError_t e;
  if (in->dtype != AST_F32) {          // kind::f16 (bf16 or fp16 operands, the format is in the instruction descriptor)
    e = set_max_smem(conv_px_kernel<0>, smem);
    if (e == cudaSuccess) launch_k(conv_px_kernel<0>, grid, PX_THREADS, smem, stream, tm_in, tm_w, p, bias, to_img32(add), to_img32(mask), to_img32(out), g->stats);
  } else {
    e = set_max_smem(conv_px_kernel<1>, smem);
    if (e == cudaSuccess) launch_k(conv_px_kernel<1>, grid, PX_THREADS, smem, stream, tm_in, tm_w, p, bias, to_img32(add), to_img32(mask), to_img32(out), g->stats);
  }
  if (e != cudaSuccess) { set_error("conv_px: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); return (int)e; }
  count_launch();
  count_work(FAM_CONV_PX, conv_flops(in, out, g), conv_bytes(in, out, g, add, mask));
  AST_CUDA_LAUNCH_CHECK();
  return 1;
}

}  // namespace ast
