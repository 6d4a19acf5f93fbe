// Shared by conv_px.cu and conv_ws.cu: the "one thread = one output channel" epilogue of the pixels-as-N kernels.
#pragma once
#include "tc_common.cuh"

namespace ast {

struct Img32 {                     // 32-bit element strides (the host checks every tensor spans < 2^31 elements)
  char* ptr;
  int dtype, h, w, sn, sh, sw;
};

__device__ __forceinline__ float px_ld(const Img32& im, int off) {
  if (im.dtype == AST_F32) return reinterpret_cast<const float*>(im.ptr)[off];
  if (im.dtype == AST_F16) return __half2float(reinterpret_cast<const __half*>(im.ptr)[off]);
  return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(im.ptr)[off]);
}
__device__ __forceinline__ void px_st(const Img32& im, int off, float v) {
  if (im.dtype == AST_F32) reinterpret_cast<float*>(im.ptr)[off] = v;
  else if (im.dtype == AST_F16) reinterpret_cast<__half*>(im.ptr)[off] = f2h_sat(v);
  else reinterpret_cast<__nv_bfloat16*>(im.ptr)[off] = __float2bfloat16_rn(v);
}

// One 32-pixel x 32-channel accumulator chunk: v[e] = D^T[ch][pixel e] for the channel `ch` this thread owns.
//   CW = 32: the 32 pixels are consecutive in x inside one tile row (conv_px, tw % 32 == 0);
//   CW = 8 : they are 4 consecutive tile rows of 8 pixels (kept for tiles of 8-pixel rows; measured slower for a
//            weight-stationary 32 x 8 variant because only cout/32 of the 4 TMEM lane quarters have epilogue warps);
//   both: offsets are affine, o(e) = base + (e / CW) * row_step + (e % CW) * col_step, valid iff e / CW < nvr and e % CW < nvc;
//   CW = 0 : lane e holds pixel e's offsets (out < 0 = invalid) and they are broadcast with shuffles.
struct PxOff { int out, add, mask; };
struct PxStep { int out_r, out_c, add_r, add_c, mask_r, mask_c; };
//   FULL: every pixel of the chunk is valid (interior tiles): the validity predicates fold away at compile time.
template <int CW, bool FULL = false>
__device__ __forceinline__ void px_chunk(float* v, const PxOff& off, const PxStep& st, int nvr, int nvc, int ch, int lane,
                                         float b, int flags, const Img32& add, const Img32& mask, const Img32& out,
                                         bool want_stats, double& s1, double& s2) {
  constexpr int W = CW ? CW : 1;
  // idx may depend on the lane (bf16 store path): ONE shuffle / one affine evaluation per use
#define PX_OFF(f, idx) (CW ? off.f + ((idx) / W) * st.f##_r + ((idx) % W) * st.f##_c : __shfl_sync(0xffffffffu, off.f, (idx)))
#define PX_VALID(idx) (FULL ? true : (CW ? ((idx) / W < nvr && (idx) % W < nvc) : __shfl_sync(0xffffffffu, off.out, (idx)) >= 0))
  // 32 operand values of this thread's channel, all loads in flight together.  The element type is tested ONCE, outside
  // the unrolled loop: a per-element type branch keeps the compiler from batching the loads (measured: the masked VGG
  // data gradients ran 4x slower when a third type joined a per-element dispatch).
#define PX_LOAD32_T(f, T, CVT, t)                                                                 \
  {                                                                                               \
    const T* base_ = reinterpret_cast<const T*>(f.ptr) + ch;                                      \
    _Pragma("unroll") for (int e = 0; e < 32; ++e) {                                              \
      const int o_ = PX_OFF(f, e);                                                                \
      t[e] = PX_VALID(e) ? CVT(base_[o_]) : 0.f;                                                  \
    }                                                                                             \
  }
#define PX_LOAD32(f, t)                                                                           \
  if (f.dtype == AST_F32) PX_LOAD32_T(f, float, float, t)                                         \
  else if (f.dtype == AST_F16) PX_LOAD32_T(f, __half, __half2float, t)                            \
  else PX_LOAD32_T(f, __nv_bfloat16, __bfloat162float, t)
  if (want_stats) {
    // InstanceNorm sums of this thread's channel over the chunk: centred on the chunk mean in fp32 (no cancellation),
    // merged into double running sums (sum x, sum x^2 = M2 + cnt * mean^2)
    // four independent partial sums: the epilogue warps are latency bound (2-4 warps per scheduler)
    float sa[4] = {0.f, 0.f, 0.f, 0.f};
    int cnt = 0;
#pragma unroll
    for (int e = 0; e < 32; ++e)
      if (PX_VALID(e)) { sa[e & 3] += v[e]; ++cnt; }
    const float sum = (sa[0] + sa[1]) + (sa[2] + sa[3]);
    if (cnt > 0) {
      const float mu = sum / (float)cnt;
      float ma[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int e = 0; e < 32; ++e)
        if (PX_VALID(e)) { const float dlt = v[e] - mu; ma[e & 3] = fmaf(dlt, dlt, ma[e & 3]); }
      const float m2 = (ma[0] + ma[1]) + (ma[2] + ma[3]);
      s1 += (double)sum;
      s2 += (double)m2 + (double)cnt * ((double)mu * (double)mu);
    }
  }
  // the add / mask operands of all 32 pixels are loaded up front (independent loads in flight), never interleaved
  // with the stores: the compiler must assume out may alias them
  if (add.ptr) {
    float t[32];
    PX_LOAD32(add, t)
#pragma unroll
    for (int e = 0; e < 32; ++e) v[e] += t[e];
  }
  // the layers under an InstanceNorm have neither bias nor activation here: skip the 64 instructions
  if (flags & AST_CONV_RELU) {
#pragma unroll
    for (int e = 0; e < 32; ++e) v[e] = fmaxf(v[e] + b, 0.f);
  } else if (b != 0.f) {
#pragma unroll
    for (int e = 0; e < 32; ++e) v[e] += b;
  }
  if (mask.ptr) {
    float t[32];
    PX_LOAD32(mask, t)
#pragma unroll
    for (int e = 0; e < 32; ++e) v[e] = t[e] > 0.f ? v[e] : 0.f;
  }
  if (flags & AST_CONV_ROUND_TF32) {
#pragma unroll
    for (int e = 0; e < 32; ++e) v[e] = round_tf32(v[e]);
  }
  if (out.dtype == AST_F32) {
#pragma unroll
    for (int e = 0; e < 32; ++e) {
      const int o = PX_OFF(out, e);
      if (FULL || (CW ? PX_VALID(e) : o >= 0)) reinterpret_cast<float*>(out.ptr)[o + ch] = v[e];
    }
  } else {
    // bf16 / fp16: neighbouring lanes trade values so that every lane stores TWO channels (4 bytes) of one pixel: even lanes
    // serve pixel e, odd lanes pixel e+1 -> 16 store instructions of 2 x 64 B instead of 32 of 64 B
    const int odd = lane & 1;
    unsigned short* const obase = reinterpret_cast<unsigned short*>(out.ptr) + (ch & ~1);   // bf16 or fp16 elements
    // CW > 0: the offset of pixel e + odd is affine in compile-time (e / CW, e % CW) plus one lane-dependent term
    const int lane_off = CW ? off.out + odd * st.out_c : 0;
#define PX_STORE16(F16)                                                                           \
    _Pragma("unroll") for (int e = 0; e < 32; e += 2) {                                           \
      const float mine = odd ? v[e + 1] : v[e];          /* my channel, the pixel I store */         \
      const float give = odd ? v[e] : v[e + 1];          /* my channel, the pixel the neighbour stores */ \
      const float got = __shfl_xor_sync(0xffffffffu, give, 1);                                    \
      const int o = CW ? lane_off + (e / W) * st.out_r + (e % W) * st.out_c : PX_OFF(out, e + odd); \
      const bool ok = FULL || (CW ? PX_VALID(e + odd) : o >= 0);                                  \
      const unsigned pk = pack2<F16>(odd ? got : mine, odd ? mine : got);                         \
      if (ok) *reinterpret_cast<unsigned*>(obase + o) = pk;                                       \
    }
    if (out.dtype == AST_F16) { PX_STORE16(true) } else { PX_STORE16(false) }
#undef PX_STORE16
  }
#undef PX_LOAD32
#undef PX_LOAD32_T
#undef PX_VALID
#undef PX_OFF
}

inline bool img32_ok(const ast_image* im) {
  if (!im) return true;
  const long long span = (long long)(im->n - 1) * im->sn + (long long)(im->h - 1) * im->sh + (long long)(im->w - 1) * im->sw + im->c;
  return im->sc == 1 && im->sn >= 0 && im->sh >= 0 && im->sw >= 0 && span < (1ll << 31);
}
inline Img32 to_img32(const ast_image* im) {
  Img32 r;
  if (!im) { r.ptr = nullptr; r.dtype = 0; r.h = r.w = r.sn = r.sh = r.sw = 0; return r; }
  r.ptr = (char*)im->ptr; r.dtype = im->dtype; r.h = im->h; r.w = im->w;
  r.sn = (int)im->sn; r.sh = (int)im->sh; r.sw = (int)im->sw;
  return r;
}


}  // namespace ast
