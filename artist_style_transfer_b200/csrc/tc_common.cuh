// PTX wrappers shared by the tcgen05 kernels (mbarrier, TMA, tcgen05.mma/ld/commit, smem descriptors).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace ast {

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Bounded spin: a protocol bug traps (launch error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  const unsigned addr = smem_u32(bar);
  for (unsigned spin = 0; spin < (1u << 26); ++spin) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    if (ok) return;
  }
  __trap();
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, unsigned long long* bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, unsigned long long* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, unsigned long long* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
template <int KIND>
__device__ __forceinline__ void tc_mma(unsigned d_tmem, unsigned long long adesc, unsigned long long bdesc,
                                       unsigned idesc, unsigned accumulate) {
  if (KIND == 0)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
  else
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld32(unsigned taddr, float* v) {
  unsigned r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// K-major shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 | LBO | SBO | version=1 | swizzle
__device__ __forceinline__ unsigned long long make_smem_desc(unsigned saddr, unsigned sbo_bytes, unsigned layout_type) {
  unsigned long long d = 0;
  d |= (unsigned long long)((saddr & 0x3FFFFu) >> 4);
  d |= (unsigned long long)1 << 16;                       // leading byte offset (unused for swizzled K-major)
  d |= (unsigned long long)(sbo_bytes >> 4) << 32;        // stride byte offset: 8 rows of one swizzle atom
  d |= (unsigned long long)1 << 46;                       // descriptor version 1 (Blackwell)
  d |= (unsigned long long)layout_type << 61;
  return d;
}
__device__ __forceinline__ unsigned long long pack_desc64(unsigned lo, unsigned hi) {
  unsigned long long d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}
__device__ __forceinline__ float round_tf32(float x) {
  unsigned r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}


// eight 4-channel vectors of one pixel, BODY applied to each (t[0..3] = channels e..e+3); element type tested once
#define LD4_EACH_T(im, off, T, BODY)                                                          \
  { _Pragma("unroll") for (int e = 0; e < 32; e += 4) { float t[4]; ld4((const T*)im.ptr + (off) + e, t); BODY } }
#define LD4_EACH(im, off, BODY)                                                               \
  if (im.dtype == AST_F32) LD4_EACH_T(im, off, float, BODY)                                   \
  else if (im.dtype == AST_F16) LD4_EACH_T(im, off, __half, BODY)                             \
  else LD4_EACH_T(im, off, __nv_bfloat16, BODY)
// 32 consecutive channels of one pixel as floats.  The element type is tested ONCE (not per vector): with a per-load type
// dispatch of three types the compiler no longer batches the loads.
__device__ __forceinline__ void ld32_img(const Img& im, long long off, float* t) {
  if (im.dtype == AST_F32) {
#pragma unroll
    for (int e = 0; e < 32; e += 4) ld4((const float*)im.ptr + off + e, t + e);
  } else if (im.dtype == AST_F16) {       // 16-bit types: four 16-byte loads (offsets are multiples of 8 elements here)
#pragma unroll
    for (int e = 0; e < 32; e += 8) {
      const uint4 u = *reinterpret_cast<const uint4*>((const __half*)im.ptr + off + e);
      const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int k = 0; k < 4; ++k) { const float2 f = __half22float2(h[k]); t[e + 2 * k] = f.x; t[e + 2 * k + 1] = f.y; }
    }
  } else {
#pragma unroll
    for (int e = 0; e < 32; e += 8) {
      const uint4 u = *reinterpret_cast<const uint4*>((const __nv_bfloat16*)im.ptr + off + e);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int k = 0; k < 4; ++k) { const float2 f = __bfloat1622float2(h[k]); t[e + 2 * k] = f.x; t[e + 2 * k + 1] = f.y; }
    }
  }
}

// Fused epilogue for 32 consecutive output channels of one pixel held in registers (conv_ws.cu):
// bias -> tap-gradient add -> ReLU -> ReLU mask -> TF32 rounding -> 16-byte stores (or scalar "thin" stores).
__device__ __forceinline__ void tc_epilogue32(float* v, int co, int img, int oy, int ox, bool thin, int cout,
                                              int cout_valid, int flags, const float* __restrict__ bias, const Img& add,
                                              const Img& mask, const Img& out) {
  if (thin) {
    // fully unrolled with a predicate: a dynamic index would force v[] (the accumulator registers) into local memory
#pragma unroll
    for (int e = 0; e < 32; ++e) {
      if (co + e >= cout_valid) continue;
      float x = v[e];
      if (bias) x += __ldg(bias + co + e);
      if (add.ptr) x += ld_elem(add, img_off(add, img, oy, ox, co + e));
      if (flags & AST_CONV_RELU) x = fmaxf(x, 0.f);
      if (mask.ptr) x = ld_elem(mask, img_off(mask, img, oy, ox, co + e)) > 0.f ? x : 0.f;
      if (flags & AST_CONV_ROUND_TF32) x = round_tf32(x);
      st_elem(out, img_off(out, img, oy, ox, co + e), x);
    }
    return;
  }
  if (co >= cout) return;
  if (bias) {
#pragma unroll
    for (int e = 0; e < 32; ++e) v[e] += __ldg(bias + co + e);
  }
  if (add.ptr) {
    const long long o = img_off(add, img, oy, ox, co);
    LD4_EACH(add, o, v[e] += t[0]; v[e + 1] += t[1]; v[e + 2] += t[2]; v[e + 3] += t[3];)
  }
  if (flags & AST_CONV_RELU) {
#pragma unroll
    for (int e = 0; e < 32; ++e) v[e] = fmaxf(v[e], 0.f);
  }
  if (mask.ptr) {
    const long long o = img_off(mask, img, oy, ox, co);
    LD4_EACH(mask, o, v[e] = t[0] > 0.f ? v[e] : 0.f; v[e + 1] = t[1] > 0.f ? v[e + 1] : 0.f; v[e + 2] = t[2] > 0.f ? v[e + 2] : 0.f; v[e + 3] = t[3] > 0.f ? v[e + 3] : 0.f;)
  }
  if (flags & AST_CONV_ROUND_TF32) {
#pragma unroll
    for (int e = 0; e < 32; ++e) v[e] = round_tf32(v[e]);
  }
  const long long oo = img_off(out, img, oy, ox, co);
  if (out.dtype == AST_F32) {
    float* op = (float*)out.ptr + oo;
#pragma unroll
    for (int e = 0; e < 32; e += 4) *reinterpret_cast<float4*>(op + e) = make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]);
  } else {
    unsigned short* op = (unsigned short*)out.ptr + oo;      // bf16 or fp16 elements
    if (out.dtype == AST_F16) {
#pragma unroll
      for (int e = 0; e < 32; e += 8)
        *reinterpret_cast<uint4*>(op + e) = make_uint4(pack2<true>(v[e], v[e + 1]), pack2<true>(v[e + 2], v[e + 3]),
                                                       pack2<true>(v[e + 4], v[e + 5]), pack2<true>(v[e + 6], v[e + 7]));
    } else {
#pragma unroll
      for (int e = 0; e < 32; e += 8)
        *reinterpret_cast<uint4*>(op + e) = make_uint4(pack2<false>(v[e], v[e + 1]), pack2<false>(v[e + 2], v[e + 3]),
                                                       pack2<false>(v[e + 4], v[e + 5]), pack2<false>(v[e + 6], v[e + 7]));
    }
  }
}

// Coalesced epilogue.  The 32 x 32 block (this warp's 32 pixel rows x 32 channels) is transposed through a 1 KB
// per-warp shared-memory stage, 8 rows per pass, XOR-swizzled 16-byte chunks (conflict-free), so every global store
// instruction writes whole 128-byte (fp32) / 64-byte (bf16) row segments.  Row offsets are exchanged ONCE per tile
// (tc_epi_row_offsets) and reused for every 32-channel chunk.  All 32 lanes must call both functions.
struct EpiRows { long long off[8]; };

// InstanceNorm statistics fused into the conv epilogue: each lane holds 32 channels of one pixel row; a reduce-scatter
// butterfly (16+8+4+2+1 shuffles per quantity) leaves lane L with the sum over the warp's 32 rows of channel co+L,
// which is added to stats[(img*C + co+L)*2 + {0,1}] (double) with one fire-and-forget atomic each.
// The two partial sums cover only this warp's 32 pixels (fp32 is accurate enough there); the accumulation across the
// plane runs in DOUBLE (per-thread running sums, atomicAdd on doubles, finalize in double), so E[x^2] - mean^2 does not
// cancel catastrophically for planes with |mean| >> std (SURVEY section 7 "shifted sums or Welford-merge").
__device__ __forceinline__ void tc_epi_stats_reduce(const float* v, bool valid, int lane, float& sum, float& sumsq) {
  float a[32], b[32];
#pragma unroll
  for (int e = 0; e < 32; ++e) { a[e] = valid ? v[e] : 0.f; b[e] = a[e] * a[e]; }
#pragma unroll
  for (int half = 16; half >= 1; half >>= 1) {
    const bool up = lane & half;
#pragma unroll
    for (int j = 0; j < half; ++j) {
      const float sa = up ? a[j] : a[j + half], ka = up ? a[j + half] : a[j];
      const float sb = up ? b[j] : b[j + half], kb = up ? b[j + half] : b[j];
      a[j] = ka + __shfl_xor_sync(0xffffffffu, sa, half);
      b[j] = kb + __shfl_xor_sync(0xffffffffu, sb, half);
    }
  }
  sum = a[0]; sumsq = b[0];
}
// conv_st.cu: persistent CTAs walk a CONTIGUOUS range of the tile list (not a grid-strided one): all tiles of a launch
// cost the same, so the ranges are balanced, and consecutive tiles of a CTA belong to the same image - the running
// InstanceNorm sums are then flushed once or twice per CTA.  (With the strided order and fewer tiles per image than CTAs
// every tile started a new image: 2 fp64 atomics per thread and tile onto 2*cout addresses cost the 32-channel 256^2
// layers ~60 us.)  Measured SLOWER for conv_hx / conv_ws (residual 3x3 44 -> 48 us, VGG conv2_2 223 -> 230 us): their
// concurrently running CTAs share halo rows and streamed weights in L2 when they work on neighbouring tiles.
__device__ __forceinline__ long long tile_begin(long long total) { return total * (long long)blockIdx.x / (long long)gridDim.x; }
__device__ __forceinline__ long long tile_end(long long total) { return total * ((long long)blockIdx.x + 1) / (long long)gridDim.x; }

// Per-thread running sums of the persistent epilogue: (sum x, sum x^2) of up to two 32-channel chunks, kept in DOUBLE across
// the tiles of one image and flushed with one atomic per quantity when the image changes (a persistent CTA walks the tiles
// of an image consecutively).  At 1080p a plane has ~16,000 tiles: flushing per tile made 33 M double atomics contend for
// 512 addresses and doubled the time of the 32-channel layers.
struct EpiStatAcc {
  double s[2][2];
  int img;
  __device__ __forceinline__ void reset(int image) { s[0][0] = s[0][1] = s[1][0] = s[1][1] = 0.0; img = image; }
  __device__ __forceinline__ void flush(double* __restrict__ stats, int cout, int co_first, int lane, int nchunks) {
    if (img < 0) return;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      if (k >= nchunks) break;
      double* row = stats + ((long long)img * cout + co_first + k * 64 + lane) * 2;
      atomicAdd(row, s[k][0]);
      atomicAdd(row + 1, s[k][1]);
    }
  }
};

__device__ __forceinline__ void sts128(unsigned addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds128(unsigned addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}

__device__ __forceinline__ void tc_epi_row_offsets(long long my_off, int lane, bool f32, EpiRows& r) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    // fp32: pass p = k>>1, iteration i = k&1, row = 8p + 4i + lane/8 ; bf16: pass p = k (k < 4), row = 8p + lane/4
    const int src = f32 ? (8 * (k >> 1) + 4 * (k & 1) + (lane >> 3)) : (8 * (k & 3) + (lane >> 2));
    r.off[k] = __shfl_sync(0xffffffffu, my_off, src);
  }
}

// 32 mask values of (pixel, channels co .. co+31): issued BEFORE the epilogue waits for the accumulator so that their
// latency hides behind the MMAs of the tile (the masked VGG dgrads spent 40 % of their time on these loads)
__device__ __forceinline__ void tc_epi_prefetch_mask(const Img& mask, int img, int oy, int ox, int co, bool valid, float* m) {
  if (!mask.ptr || !valid) return;
  const long long o = img_off(mask, img, oy, ox, co);
  ld32_img(mask, o, m);
}

__device__ __forceinline__ void tc_epilogue32_coalesced(float* v, int co, int img, int oy, int ox, bool valid, int cout,
                                                        int flags, const float* __restrict__ bias, const Img& add,
                                                        const Img& mask, const Img& out, const EpiRows& rows,
                                                        unsigned char* stage, int lane, const float* pre_mask, bool have_pre) {
  if (co >= cout) return;                                   // uniform
  if (valid) {
    if (bias) {
#pragma unroll
      for (int e = 0; e < 32; e += 4) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + co + e));
        v[e] += b4.x; v[e + 1] += b4.y; v[e + 2] += b4.z; v[e + 3] += b4.w;
      }
    }
    if (add.ptr) {
      const long long o = img_off(add, img, oy, ox, co);
      LD4_EACH(add, o, v[e] += t[0]; v[e + 1] += t[1]; v[e + 2] += t[2]; v[e + 3] += t[3];)
    }
    if (flags & AST_CONV_RELU) {
#pragma unroll
      for (int e = 0; e < 32; ++e) v[e] = fmaxf(v[e], 0.f);
    }
    if (mask.ptr && have_pre) {
#pragma unroll
      for (int e = 0; e < 32; ++e) v[e] = pre_mask[e] > 0.f ? v[e] : 0.f;
    } else if (mask.ptr) {
      const long long o = img_off(mask, img, oy, ox, co);
      LD4_EACH(mask, o, v[e] = t[0] > 0.f ? v[e] : 0.f; v[e + 1] = t[1] > 0.f ? v[e + 1] : 0.f; v[e + 2] = t[2] > 0.f ? v[e + 2] : 0.f; v[e + 3] = t[3] > 0.f ? v[e + 3] : 0.f;)
    }
    if (flags & AST_CONV_ROUND_TF32) {
#pragma unroll
      for (int e = 0; e < 32; ++e) v[e] = round_tf32(v[e]);
    }
  }
  const unsigned st = smem_u32(stage);                      // explicit shared-space accesses (16-byte slots)
  const int r8 = lane & 7;
  if (out.dtype == AST_F32) {                               // stage: [8 rows][8 chunks of 16 B]
    float* base = (float*)out.ptr + co + 4 * (lane & 7);
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      if ((lane >> 3) == p) {
#pragma unroll
        for (int c = 0; c < 8; ++c)
          sts128(st + 16u * (r8 * 8 + (c ^ r8)), make_uint4(__float_as_uint(v[4 * c]), __float_as_uint(v[4 * c + 1]),
                                                             __float_as_uint(v[4 * c + 2]), __float_as_uint(v[4 * c + 3])));
      }
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int row = (lane >> 3) + 4 * i;
        const uint4 val = lds128(st + 16u * (row * 8 + ((lane & 7) ^ row)));
        const long long off = rows.off[2 * p + i];
        if (off >= 0) *reinterpret_cast<uint4*>(base + off) = val;
      }
      __syncwarp();
    }
  } else {                                                  // stage: [8 rows][4 chunks of 16 B]
    unsigned short* base = (unsigned short*)out.ptr + co + 8 * (lane & 3);      // bf16 or fp16 elements
    unsigned w[16];                                          // this lane's 32 channels as 16-bit pairs; type tested once
    if (out.dtype == AST_F16) {
#pragma unroll
      for (int i = 0; i < 16; ++i) w[i] = pack2<true>(v[2 * i], v[2 * i + 1]);
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) w[i] = pack2<false>(v[2 * i], v[2 * i + 1]);
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      if ((lane >> 3) == p) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
          sts128(st + 16u * (r8 * 4 + (c ^ (r8 & 3))), make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]));
      }
      __syncwarp();
      {
        const int row = lane >> 2;
        const uint4 val = lds128(st + 16u * (row * 4 + ((lane & 3) ^ (row & 3))));
        const long long off = rows.off[p];
        if (off >= 0) *reinterpret_cast<uint4*>(base + off) = val;
      }
      __syncwarp();
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// cuTensorMapEncodeTiled costs a few microseconds on the host and ran twice per launch; the maps only depend on
// (pointer, shape, strides, box), which repeat from step to step (PyTorch's caching allocator hands the same blocks
// out again), so they are cached.  The map is copied out by value (128 bytes): no pointer into the cache escapes.
int cached_tensor_map(EncodeTiledFn encode, CUtensorMap* out, CUtensorMapDataType dt, int rank, void* ptr,
                      const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box, const cuuint32_t* estr,
                      CUtensorMapSwizzle sw, CUtensorMapL2promotion l2);
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) only when the request grows (per device and kernel)
cudaError_t set_max_smem_impl(const void* kernel, size_t smem);
template <typename K> inline cudaError_t set_max_smem(K kernel, size_t smem) { return set_max_smem_impl((const void*)kernel, smem); }

// power-of-two (w x h = area) pixel box that wastes the fewest positions of an mi x mj grid
// operand dtype -> TMA element type / tcgen05 operand format (kind::tf32 for fp32 storage, kind::f16 with bf16 or fp16)
inline CUtensorMapDataType tc_tmap_dtype(int dtype) {
  return dtype == AST_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : (dtype == AST_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
}
inline unsigned tc_operand_fmt(int dtype) { return dtype == AST_F32 ? 2u : (dtype == AST_F16 ? 0u : 1u); }

inline void pick_tile(int mi, int mj, int area, int* tw, int* th) {
  double best = -1;
  for (int w = area; w >= 1; w >>= 1) {
    const int h = area / w;
    if (w > 256 || h > 256) continue;
    const double cover = (double)((mi + h - 1) / h * h) * ((mj + w - 1) / w * w);
    const double eff = (double)mi * mj / cover;
    if (eff > best + 1e-9) { best = eff; *tw = w; *th = h; }
  }
}

}  // namespace ast
