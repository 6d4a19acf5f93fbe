// tcgen05 / TMEM / TMA implicit-GEMM "gather" convolution for sm_100a (Blackwell B200).
//
//   D[128 pixels x BN couts] (fp32, TMEM)  +=  A_tap[128 pixels x KC cin] (smem, TMA)  x  W_tap[BN x KC]^T (smem, TMA)
//
// One persistent CTA per SM, warp-specialised (DESIGN.md "conv_tc"):
//   warp 0      TMA producers (up to 4 lanes take the stages round-robin): per (tap, cin-chunk) one 4-D box of the
//               NHWC activation tensor (OOB = zero padding, elementStrides = input stride) and one 2-D box of the
//               packed weights, 128B/64B-swizzled, K-major
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (kind::f16 for bf16, kind::tf32 for fp32 data),
//               tcgen05.commit releases smem stages / publishes the accumulator
//   warps 2-9   epilogue (two warps per TMEM lane quarter): tcgen05.ld 32x32b -> fused InstanceNorm sums / bias /
//               tap-gradient add / ReLU / ReLU-mask / tf32 rounding -> smem-transposed 16-byte stores to the NHWC
//               output (double-buffered accumulator in TMEM)
// This file is the dispatcher of the tensor-core convolutions as well: conv_gather_tc() first offers a launch to the
// weight-stationary kernel (conv_ws.cu) and to the pixels-as-N kernel (conv_px.cu, cout 64/128) and runs the kernel below
// for everything else (256/512-channel layers, short K loops, 3-channel "thin" outputs, per-image weights).
// Replaces cuDNN/oneDNN convolution forward / backward-data at the call sites listed in include/ast.h.
#include "tc_common.cuh"

namespace ast {

constexpr int TC_THREADS = 320;   // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (2 per TMEM lane quarter)
constexpr int MAX_STAGES = 8;
#ifndef TC_PRODUCERS
#define TC_PRODUCERS 4
#endif

struct TcParams {
  int mi, mj, tw, th, tiles_i, tiles_j, n_img, n_tiles_n, bn;
  int ntaps, kchunks, kc;
  int si, so, oy0, ox0;
  int cout, cout_valid, thin, flags;   // cout = weight rows per tap (multiple of 32); thin: cout_valid < cout or strided out
  int w_rows_per_img;
  int stages, a_bytes, stage_bytes, rowb;
  unsigned idesc, layout_type, sbo;
  long long total_tiles;
  short dy[AST_MAX_TAPS];
  short dx[AST_MAX_TAPS];
};

// ------------------------------------------------------------------ kernel
// (A cta_group::2 form - CTA pairs, M = 256, half of the weight tile per CTA - was measured to give no gain on this path
//  and removed: scratch/dead_variants/conv_tc_r01.cu.txt, profiles/r01_summary.md finding 4.)
template <int KIND>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_w, const TcParams p,
               const float* __restrict__ bias, const Img add, const Img mask, const Img out,
               double* __restrict__ stats) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long full_bar[MAX_STAGES], empty_bar[MAX_STAGES], tfull_bar[2], tempty_bar[2];
  __shared__ unsigned tmem_slot;

  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ksteps = p.ntaps * p.kchunks;
  const unsigned tmem_cols = (2 * p.bn <= 32) ? 32 : (2 * p.bn <= 64) ? 64 : (2 * p.bn <= 128) ? 128 : (2 * p.bn <= 256) ? 256 : 512;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_in) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_w) : "memory");
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 256); }
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

  if (warp == 0) {
    // ============================ TMA producer ============================
    // TC_PRODUCERS lanes take the pipeline stages round-robin: ONE issuing thread sustains only ~45 B/clk (a wait +
    // expect_tx + two TMA instructions cost it ~750 cycles), four reach the ~80 B/clk an SM can pull from L2
    // (scratch/mma_bench.cu, profiles/r01_summary.md).
    // (at most `stages` lanes: a lane running two ring cycles ahead would alias the 1-bit mbarrier parity)
    int s = 0, turn = 0; unsigned ph = 0;
    const int nprod = p.stages < TC_PRODUCERS ? p.stages : TC_PRODUCERS;
    if (lane < nprod)
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      long long r = tile;
      const int nt = (int)(r % p.n_tiles_n); r /= p.n_tiles_n;
      const int tj = (int)(r % p.tiles_j); r /= p.tiles_j;
      const int ti = (int)(r % p.tiles_i);
      const int img = (int)(r / p.tiles_i);
      const int x0 = p.si * tj * p.tw, y0 = p.si * ti * p.th;
      const int wrow0 = img * p.w_rows_per_img + nt * p.bn;
      for (int t = 0; t < p.ntaps; ++t) {
        const int cx = x0 + p.dx[t], cy = y0 + p.dy[t];
        for (int kc = 0; kc < p.kchunks; ++kc) {
          if (turn == lane) {
            mbar_wait(&empty_bar[s], ph ^ 1);
            unsigned char* sa = smem + (size_t)s * p.stage_bytes;
            mbar_expect_tx(&full_bar[s], (unsigned)p.stage_bytes);
            tma_load_4d(sa, &tm_in, &full_bar[s], kc * p.kc, cx, cy, img);
            tma_load_2d(sa + p.a_bytes, &tm_w, &full_bar[s], kc * p.kc, wrow0 + t * p.cout);
          }
          if (++turn == nprod) turn = 0;
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    int s = 0; unsigned ph = 0; int as = 0; unsigned aph = 0;
    const int kmma = p.rowb / 32;   // UMMA_K spans 32 bytes for both bf16 (16 elems) and tf32 (8 elems)
    const unsigned desc_hi = (p.sbo >> 4) | (1u << 14) | (p.layout_type << 29);
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      mbar_wait(&tempty_bar[as], aph ^ 1);
      tc_fence_after();
      const unsigned d_tmem = tmem_base + (unsigned)(as * p.bn);
      for (int ks = 0; ks < ksteps; ++ks) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        if (lane == 0) {
          // descriptor hi word is constant; lo word = (addr >> 4) | LBO, advanced by 32 B (= 2) per UMMA_K step
          const unsigned a_addr = smem_u32(smem + (size_t)s * p.stage_bytes);
          const unsigned a_lo = ((a_addr & 0x3FFFFu) >> 4) | (1u << 16);
          const unsigned b_lo = (((a_addr + p.a_bytes) & 0x3FFFFu) >> 4) | (1u << 16);
          tc_mma<KIND>(d_tmem, pack_desc64(a_lo, desc_hi), pack_desc64(b_lo, desc_hi), p.idesc, ks > 0 ? 1u : 0u);
          tc_mma<KIND>(d_tmem, pack_desc64(a_lo + 2, desc_hi), pack_desc64(b_lo + 2, desc_hi), p.idesc, 1u);
          if (kmma == 4) {
            tc_mma<KIND>(d_tmem, pack_desc64(a_lo + 4, desc_hi), pack_desc64(b_lo + 4, desc_hi), p.idesc, 1u);
            tc_mma<KIND>(d_tmem, pack_desc64(a_lo + 6, desc_hi), pack_desc64(b_lo + 6, desc_hi), p.idesc, 1u);
          }
          tc_commit(&empty_bar[s]);                 // frees the smem stage when these MMAs retire
          if (ks == ksteps - 1) tc_commit(&tfull_bar[as]);   // accumulator complete -> epilogue
        }
        __syncwarp();
        if (++s == p.stages) { s = 0; ph ^= 1; }
      }
      if (++as == 2) { as = 0; aph ^= 1; }
    }
  } else if (warp >= 2) {
    // ============================ epilogue (warps 2..9) ============================
    const int q = warp & 3;                        // TMEM lane quarter this warp may access
    const int cpar = (warp - 2) >> 2;              // which half of the 32-column chunks this warp drains
    unsigned char* stage = smem + (size_t)p.stages * p.stage_bytes + (warp - 2) * 1024;
    const int row = q * 32 + lane;
    const int ty = row / p.tw, tx = row % p.tw;
    int as = 0; unsigned aph = 0;
    EpiStatAcc sacc;
    sacc.reset(-1);
    const int s_chunks = (p.bn - cpar * 32 + 63) / 64;          // 32-channel chunks this warp drains per tile
    const bool s_run = stats && s_chunks <= 2 && p.n_tiles_n == 1;
    for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      long long r = tile;
      const int nt = (int)(r % p.n_tiles_n); r /= p.n_tiles_n;
      const int tj = (int)(r % p.tiles_j); r /= p.tiles_j;
      const int ti = (int)(r % p.tiles_i);
      const int img = (int)(r / p.tiles_i);
      if (s_run && img != sacc.img) { sacc.flush(stats, p.cout_valid, cpar * 32, lane, s_chunks); sacc.reset(img); }
      const int i = ti * p.th + ty, j = tj * p.tw + tx;
      const int oy = p.oy0 + p.so * i, ox = p.ox0 + p.so * j;
      const bool valid = i < p.mi && j < p.mj && oy < out.h && ox < out.w;
      mbar_wait(&tfull_bar[as], aph);
      tc_fence_after();
      const unsigned taddr0 = tmem_base + ((unsigned)(q * 32) << 16) + (unsigned)(as * p.bn);
      EpiRows rows;
      if (!p.thin) tc_epi_row_offsets(valid ? img_off(out, img, oy, ox, 0) : -1, lane, out.dtype == AST_F32, rows);
      for (int c0 = cpar * 32; c0 < p.bn; c0 += 64) {
        float v[32];
        tc_ld32(taddr0 + c0, v);
        const int co = nt * p.bn + c0;
        if (p.thin) {
          if (valid) tc_epilogue32(v, co, img, oy, ox, true, p.cout, p.cout_valid, p.flags, bias, add, mask, out);
        } else {
          if (stats) {
            float su, sq;
            tc_epi_stats_reduce(v, valid, lane, su, sq);
            if (s_run) { const int k = (c0 - cpar * 32) >> 6; sacc.s[k & 1][0] += (double)su; sacc.s[k & 1][1] += (double)sq; }
            else { double* row = stats + ((long long)img * p.cout_valid + co + lane) * 2; atomicAdd(row, (double)su); atomicAdd(row + 1, (double)sq); }
          }
          // (a mask prefetch like conv_ws's was tried here: the extra registers slowed the unmasked layers more than
          // the masked deep VGG dgrads gained)
          tc_epilogue32_coalesced(v, co, img, oy, ox, valid, p.cout, p.flags, bias, add, mask, out, rows, stage, lane,
                                  v, false);
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

// ------------------------------------------------------------------ host side
int conv_gather_ws(const ast_image* in, const void* weights, const float* bias, const ast_image* add,
                   const ast_image* mask, const ast_image* out, const ast_gather_geom* g, int cpad, bool thin,
                   cudaStream_t stream);   // conv_ws.cu: 1 = launched, 0 = not applicable
int conv_gather_px(const ast_image* in, const void* weights, const float* bias, const ast_image* add,
                   const ast_image* mask, const ast_image* out, const ast_gather_geom* g, int cpad, bool thin,
                   cudaStream_t stream);   // conv_px.cu: 1 = launched, 0 = not applicable
int conv_gather_hx(const ast_image* in, const void* weights, const float* bias, const ast_image* add,
                   const ast_image* mask, const ast_image* out, const ast_gather_geom* g, int cpad, bool thin,
                   cudaStream_t stream);   // conv_hx.cu: 1 = launched, 0 = not applicable
int tc_capabilities() { return 3; }   // 1 = conv_tc.cu, 2 = contract_tc.cu (both are always built together)

int conv_gather_tc(const ast_image* in, const void* weights, const float* bias, const float* in_shift,
                   const ast_image* add, const ast_image* mask, const ast_image* out, const ast_gather_geom* g,
                   cudaStream_t stream) {
  const int esz = in->dtype == AST_F32 ? 4 : 2;
  AST_CHECK_ARG(!in_shift, "conv_tc: in_shift is only supported by the SIMT kernel");
  AST_CHECK_ARG(!(g->flags & AST_CONV_REFLECT), "conv_tc: reflect addressing needs a physically padded input");
  const int cpad = (out->c + 31) / 32 * 32;        // weight rows per tap the caller packed (zero rows beyond out->c)
  const bool thin = cpad != out->c || out->sc != 1;
  AST_CHECK_ARG(in->sc == 1, "conv_tc: NHWC input required");
  AST_CHECK_ARG(!g->stats || !thin, "conv_tc: fused statistics need a full NHWC output (cout %% 32 == 0)");
  AST_CHECK_ARG((in->c * esz) % 64 == 0, "conv_tc: cin*elemsize must be a multiple of 64 bytes (cin=%d)", in->c);
  AST_CHECK_ARG(g->si >= 1 && g->si <= 2, "conv_tc: input stride must be 1 or 2");
  AST_CHECK_ARG(((uintptr_t)in->ptr & 15) == 0 && ((uintptr_t)weights & 15) == 0 && (thin || ((uintptr_t)out->ptr & 15) == 0),
                "conv_tc: pointers must be 16-byte aligned");
  AST_CHECK_ARG((in->sw * esz) % 16 == 0 && (in->sh * esz) % 16 == 0 && (in->sn * esz) % 16 == 0,
                "conv_tc: input strides must be multiples of 16 bytes");
  AST_CHECK_ARG(thin || (out->sw % 4 == 0 && out->sh % 4 == 0 && out->sn % 4 == 0), "conv_tc: output strides must be multiples of 4 elements");
  AST_CHECK_ARG(thin || !add || (add->sc == 1 && add->sw % 4 == 0 && add->sh % 4 == 0 && add->sn % 4 == 0), "conv_tc: add layout");
  AST_CHECK_ARG(thin || !mask || (mask->sc == 1 && mask->sw % 4 == 0 && mask->sh % 4 == 0 && mask->sn % 4 == 0), "conv_tc: mask layout");
  // 16-bit add / mask operands are read with 16-byte loads
  AST_CHECK_ARG(thin || !mask || mask->dtype == AST_F32 || (mask->sw % 8 == 0 && mask->sh % 8 == 0 && mask->sn % 8 == 0 && ((uintptr_t)mask->ptr & 15) == 0), "conv_tc: 16-bit mask strides must be multiples of 8 elements");
  AST_CHECK_ARG(thin || !add || add->dtype == AST_F32 || (add->sw % 8 == 0 && add->sh % 8 == 0 && add->sn % 8 == 0 && ((uintptr_t)add->ptr & 15) == 0), "conv_tc: 16-bit add strides must be multiples of 8 elements");
  if (in->n == 0) return 0;
  EncodeTiledFn encode = get_encode();
  AST_CHECK_ARG(encode, "conv_tc: cuTensorMapEncodeTiled entry point not available");
  // dispatch: weight-stationary kernel (filter resident in smem, stride 1), then the halo-reusing pixels-as-N kernel
  // (stride 1, cout % 128 == 0, multi-tap), then the per-tap pixels-as-N kernel (cout 64/128, long K loops: stride-2
  // layers), then the generic kernel below (short K loops, per-image weights, thin outputs)
  if (int wr = conv_gather_ws(in, weights, bias, add, mask, out, g, cpad, thin, stream)) return wr == 1 ? 0 : wr;
  AST_CHECK_ARG(!g->pooled, "conv_tc: a pooled output needs the weight-stationary kernel (stride 1, filter resident in smem)");
  if (int hr = conv_gather_hx(in, weights, bias, add, mask, out, g, cpad, thin, stream)) return hr == 1 ? 0 : hr;
  if (int pr = conv_gather_px(in, weights, bias, add, mask, out, g, cpad, thin, stream)) return pr == 1 ? 0 : pr;
  AST_CHECK_ARG(in->dtype != AST_F16, "conv_tc: fp16 operands are supported by the weight-stationary and halo kernels only (stride 1)");

  TcParams p;
  memset(&p, 0, sizeof(p));
  p.mi = g->mi; p.mj = g->mj; p.si = g->si; p.so = g->so; p.oy0 = g->oy0; p.ox0 = g->ox0;
  p.ntaps = g->ntaps; p.flags = g->flags; p.cout = cpad; p.cout_valid = out->c; p.thin = thin; p.n_img = in->n;
  for (int t = 0; t < g->ntaps; ++t) { p.dy[t] = g->dy[t]; p.dx[t] = g->dx[t]; }
  pick_tile(p.mi, p.mj, 128, &p.tw, &p.th);
  p.tiles_i = (p.mi + p.th - 1) / p.th;
  p.tiles_j = (p.mj + p.tw - 1) / p.tw;
  p.rowb = (in->c * esz) % 128 == 0 ? 128 : 64;
  p.kc = p.rowb / esz;
  p.kchunks = in->c / p.kc;
  p.bn = cpad % 256 == 0 ? 256 : (cpad <= 256 ? cpad : (cpad % 128 == 0 ? 128 : (cpad % 64 == 0 ? 64 : 32)));
  AST_CHECK_ARG(p.bn == 32 || p.bn == 64 || p.bn == 128 || p.bn == 256, "conv_tc: unsupported cout %d", out->c);
  p.n_tiles_n = cpad / p.bn;
  p.w_rows_per_img = 0;
  if (g->w_img_stride) {
    AST_CHECK_ARG(g->w_img_stride == (int64_t)g->ntaps * cpad * in->c, "conv_tc: per-image weights must be densely packed");
    p.w_rows_per_img = g->ntaps * cpad;
  }
  p.a_bytes = 128 * p.rowb;
  p.stage_bytes = p.a_bytes + p.bn * p.rowb;
  p.stages = (200 * 1024) / p.stage_bytes;
  if (p.stages > MAX_STAGES) p.stages = MAX_STAGES;
  p.layout_type = p.rowb == 128 ? 2u : 4u;       // SWIZZLE_128B / SWIZZLE_64B
  p.sbo = 8u * p.rowb;
  const unsigned fmt = tc_operand_fmt(in->dtype);
  p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((unsigned)(p.bn >> 3) << 17) | ((128u >> 4) << 24);
  p.total_tiles = (long long)p.n_img * p.tiles_i * p.tiles_j * p.n_tiles_n;

  // ---- tensor maps
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
    cuuint32_t box[2] = {(cuuint32_t)p.kc, (cuuint32_t)p.bn};
    cuuint32_t estr[2] = {1, 1};
    if (int r = cached_tensor_map(encode, &tm_w, dt, 2, const_cast<void*>(weights), dims, strides, box, estr, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B)) return r;
  }
  const size_t smem = (size_t)p.stages * p.stage_bytes + 1024 + 8192;   // + per-warp store-transpose stage
  const int grid = (int)(p.total_tiles < num_sms() ? p.total_tiles : num_sms());
  Img addi = add ? to_img(add) : null_img(), maski = mask ? to_img(mask) : null_img();
  cudaError_t e;
  if (in->dtype != AST_F32) {          // kind::f16 (bf16 or fp16 operands, the format is in the instruction descriptor)
    e = set_max_smem(conv_tc_kernel<0>, smem);
    if (e == cudaSuccess) launch_k(conv_tc_kernel<0>, grid, TC_THREADS, smem, stream, tm_in, tm_w, p, bias, addi, maski, to_img(out), g->stats);
  } else {
    e = set_max_smem(conv_tc_kernel<1>, smem);
    if (e == cudaSuccess) launch_k(conv_tc_kernel<1>, grid, TC_THREADS, smem, stream, tm_in, tm_w, p, bias, addi, maski, to_img(out), g->stats);
  }
  if (e != cudaSuccess) { set_error("conv_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); return (int)e; }
  count_launch();
  count_work(FAM_CONV_TC, conv_flops(in, out, g), conv_bytes(in, out, g, add, mask));
  AST_CUDA_LAUNCH_CHECK();
  return 0;
}

}  // namespace ast
