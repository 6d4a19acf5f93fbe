// Shared device/host helpers for libast_b200.so (sm_100a only).
#pragma once
#include <cstring>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/ast.h"

namespace ast {

// ---- error plumbing (no C++ exceptions cross the ABI; SURVEY 8b "Errors") ----
void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define AST_CHECK_ARG(cond, ...)                    \
  do {                                              \
    if (!(cond)) {                                  \
      ast::set_error(__VA_ARGS__);                  \
      return -1;                                    \
    }                                               \
  } while (0)

#define AST_CUDA_LAUNCH_CHECK()                                             \
  do {                                                                      \
    cudaError_t e__ = cudaGetLastError();                                   \
    if (e__ != cudaSuccess) {                                               \
      ast::set_error("%s:%d launch failed: %s", __FILE__, __LINE__,         \
                     cudaGetErrorString(e__));                              \
      return (int)e__;                                                      \
    }                                                                       \
  } while (0)

// ---- programmatic dependent launch (PDL) -------------------------------------------------------------------------
// Every kernel of the library is launched through launch_k().  With AST_PDL=1 the launch carries
// cudaLaunchAttributeProgrammaticStreamSerialization: the grid may be scheduled while its predecessor in the stream is
// still draining, and every kernel starts with pdl_sync() = { griddepcontrol.launch_dependents; griddepcontrol.wait; }:
// it lets ITS successor be scheduled early and then blocks until the predecessor grid has completed and its memory is
// visible.  No kernel touches global memory before pdl_sync(), so the semantics are those of plain stream order.
// Both instructions are no-ops without the attribute.  Measured on the B=32 step (CUDA graph): 13.86 ms with PDL vs
// 13.55 ms without when every kernel triggers early (successors take SM slots from multi-wave kernels), 13.65 ms when only
// the single-wave persistent kernels trigger - so it is OFF by default.
__device__ __forceinline__ void pdl_sync() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// Only the single-wave persistent kernels let their successor be scheduled early (its CTAs then fill SMs as this grid's
// CTAs retire); multi-wave kernels would lose SM slots to the waiting successor.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
inline bool pdl_enabled() {
  static const int on = [] { const char* e = getenv("AST_PDL"); return e ? atoi(e) : 0; }();
  return on != 0;
}
template <typename... KP, typename... A>
inline cudaError_t launch_k(void (*kernel)(KP...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, A&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KP>(args)...);
}

// ---- device-side image view ----
struct Img {
  char* ptr;
  int dtype;
  int n, h, w, c;
  long long sn, sh, sw, sc;
};

inline Img to_img(const ast_image* a) {
  Img r;
  r.ptr = (char*)a->ptr; r.dtype = a->dtype; r.n = a->n; r.h = a->h; r.w = a->w; r.c = a->c;
  r.sn = a->sn; r.sh = a->sh; r.sw = a->sw; r.sc = a->sc;
  return r;
}
inline Img null_img() { Img r; r.ptr = nullptr; r.dtype = 0; r.n = r.h = r.w = r.c = 0; r.sn = r.sh = r.sw = r.sc = 0; return r; }
inline bool same_shape(const ast_image* a, const ast_image* b) {
  return a->n == b->n && a->h == b->h && a->w == b->w && a->c == b->c;
}

template <typename T> struct DT;
template <> struct DT<float> {
  static constexpr int id = AST_F32;
  __device__ static __forceinline__ float ld(const float* p) { return *p; }
  __device__ static __forceinline__ void st(float* p, float v) { *p = v; }
};
template <> struct DT<__nv_bfloat16> {
  static constexpr int id = AST_BF16;
  __device__ static __forceinline__ float ld(const __nv_bfloat16* p) { return __bfloat162float(*p); }
  __device__ static __forceinline__ void st(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};

__device__ __forceinline__ float ld_elem(const Img& im, long long off) {
  return im.dtype == AST_F32 ? ((const float*)im.ptr)[off] : __bfloat162float(((const __nv_bfloat16*)im.ptr)[off]);
}
__device__ __forceinline__ void st_elem(const Img& im, long long off, float v) {
  if (im.dtype == AST_F32) ((float*)im.ptr)[off] = v;
  else ((__nv_bfloat16*)im.ptr)[off] = __float2bfloat16_rn(v);
}
__device__ __forceinline__ long long img_off(const Img& im, int n, int y, int x, int c) {
  return (long long)n * im.sn + (long long)y * im.sh + (long long)x * im.sw + (long long)c * im.sc;
}
// nn.ReflectionPad2d index rule (no edge repeat): -1 -> 1, H -> H-2
__device__ __forceinline__ int reflect_idx(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// 16-byte vector of elements (4 fp32 / 8 bf16) unpacked to floats
template <typename T> struct Vec16;
template <> struct Vec16<float> {
  static constexpr int N = 4;
  __device__ static __forceinline__ void load(const float* p, float* out) {
    float4 v = *reinterpret_cast<const float4*>(p);
    out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
  }
  __device__ static __forceinline__ void store(float* p, const float* in) {
    *reinterpret_cast<float4*>(p) = make_float4(in[0], in[1], in[2], in[3]);
  }
};
template <> struct Vec16<__nv_bfloat16> {
  static constexpr int N = 8;
  __device__ static __forceinline__ void load(const __nv_bfloat16* p, float* out) {
    uint4 v = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); out[2 * i] = f.x; out[2 * i + 1] = f.y; }
  }
  __device__ static __forceinline__ void store(__nv_bfloat16* p, const float* in) {
    uint4 v;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(in[2 * i], in[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = v;
  }
};

// 4 consecutive channels (16 B fp32 / 8 B bf16); offsets must be multiples of 4 elements
__device__ __forceinline__ void ld4(const float* p, float* v) {
  const float4 f = *reinterpret_cast<const float4*>(p);
  v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
}
__device__ __forceinline__ void ld4(const __nv_bfloat16* p, float* v) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
  const float2 f0 = __bfloat1622float2(h[0]), f1 = __bfloat1622float2(h[1]);
  v[0] = f0.x; v[1] = f0.y; v[2] = f1.x; v[3] = f1.y;
}
__device__ __forceinline__ void st4(float* p, const float* v) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void st4(__nv_bfloat16* p, const float* v) {
  uint2 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
  h[0] = __floats2bfloat162_rn(v[0], v[1]);
  h[1] = __floats2bfloat162_rn(v[2], v[3]);
  *reinterpret_cast<uint2*>(p) = u;
}
__device__ __forceinline__ void ld4_img(const Img& im, long long off, float* v) {
  if (im.dtype == AST_F32) ld4((const float*)im.ptr + off, v);
  else ld4((const __nv_bfloat16*)im.ptr + off, v);
}
__device__ __forceinline__ void st4_img(const Img& im, long long off, const float* v) {
  if (im.dtype == AST_F32) st4((float*)im.ptr + off, v);
  else st4((__nv_bfloat16*)im.ptr + off, v);
}

inline int num_sms() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

}  // namespace ast
