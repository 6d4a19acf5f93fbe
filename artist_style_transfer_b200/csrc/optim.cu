// Optimizer side of the step (train_cnn.py:247-248,334,375) as ONE table-driven launch over a flat parameter arena:
//   g <- G (kernel-native gradient layout, read through a per-tensor affine map) + weight_decay * p      (Adam's L2 term)
//   m, v moving averages, bias correction, p <- p - lr * mhat / (sqrt(vhat) + eps)
//   and the refreshed bf16 / TF32 / fp32 operand copies ("packs") the next step's convolutions read, in their
//   [tap][cout][cin] layouts - so the 35 per-layer pack launches, torch's fused Adam and the gradient flatten /
//   unflatten copies around the all-reduce disappear.
// lr / step / bias corrections live in DEVICE memory (ast_adam_state): a captured CUDA graph replays with the values
// of the moment, so StepLR (train_cnn.py:248,375) keeps working under graph replay.
// Replaces torch.optim.Adam's kernels, aten::_foreach_copy_, aten::sum (dbeta/dgamma/bias reductions).
#include "common.cuh"

namespace ast {

constexpr int OPT_THREADS = 256;
constexpr int OPT_PER_THREAD = 4;
constexpr int OPT_ITEM = OPT_THREADS * OPT_PER_THREAD;

__global__ void adam_tick_kernel(ast_adam_state* st) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const float t = st->step + 1.f;
    st->step = t;
    st->bias_c1 = (float)(1.0 - pow((double)st->beta1, (double)t));
    st->bias_c2 = (float)(1.0 - pow((double)st->beta2, (double)t));
  }
}

__device__ __forceinline__ void store_pack(void* arena, long long byte_off, long long elem, int dtype, float v) {
  char* base = reinterpret_cast<char*>(arena) + byte_off;
  if (dtype == AST_BF16) reinterpret_cast<__nv_bfloat16*>(base)[elem] = __float2bfloat16_rn(v);
  else if (dtype == AST_TF32) {
    unsigned r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    reinterpret_cast<float*>(base)[elem] = __uint_as_float(r);
  } else reinterpret_cast<float*>(base)[elem] = v;
}

template <bool UPDATE>
__global__ void __launch_bounds__(OPT_THREADS)
adam_pack_kernel(const ast_param_desc* __restrict__ descs, const int32_t* __restrict__ work, float* __restrict__ params,
                 const float* __restrict__ grads, float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq,
                 void* __restrict__ pack_arena, const int64_t* __restrict__ taps, const ast_adam_state* __restrict__ st) {
  __shared__ ast_param_desc d;
  const int item = blockIdx.x;
  {
    const int* src = reinterpret_cast<const int*>(descs + work[2 * item]);
    int* dst = reinterpret_cast<int*>(&d);
    for (int k = threadIdx.x; k < (int)(sizeof(ast_param_desc) / sizeof(int)); k += OPT_THREADS) dst[k] = src[k];
  }
  __syncthreads();
  const long long start = (long long)work[2 * item + 1];
  float lr = 0.f, b1 = 0.f, b2 = 0.f, eps = 0.f, wd = 0.f, gs = 1.f, step_size = 0.f, inv_sqrt_bc2 = 1.f;
  if (UPDATE) {
    lr = st->lr; b1 = st->beta1; b2 = st->beta2; eps = st->eps; wd = st->weight_decay; gs = st->grad_scale;
    step_size = lr / st->bias_c1;
    inv_sqrt_bc2 = rsqrtf(st->bias_c2);
  }
  const int B = d.dim[1], U = d.dim[2], V = d.dim[3];
  const int UV = U * V;
#pragma unroll
  for (int e = 0; e < OPT_PER_THREAD; ++e) {
    const long long i = start + (long long)e * OPT_THREADS + threadIdx.x;     // coalesced over the thread index
    if (i >= d.numel) continue;
    const int uv = (int)(i % UV);
    const long long ab = i / UV;
    const int b = (int)(ab % B);
    const long long a = ab / B;
    float p = params[d.p_off + i];
    if (UPDATE) {
      const float g = fmaf(wd, p, gs * __ldg(grads + d.g_off + a * d.g_stride[0] + b * d.g_stride[1] + taps[d.g_tap + uv]));
      const float m = fmaf(b1, exp_avg[d.s_off + i], (1.f - b1) * g);
      const float v = fmaf(b2, exp_avg_sq[d.s_off + i], (1.f - b2) * g * g);
      exp_avg[d.s_off + i] = m;
      exp_avg_sq[d.s_off + i] = v;
      p -= step_size * m / (sqrtf(v) * inv_sqrt_bc2 + eps);
      params[d.p_off + i] = p;
    }
#pragma unroll
    for (int k = 0; k < 2; ++k)
      if (k < d.n_pack) {
        const long long at = a * d.pack[k].stride[0] + b * d.pack[k].stride[1] + taps[d.pack[k].tap + uv];
        for (int r = 0; r < d.pack[k].rep; ++r)
          store_pack(pack_arena, d.pack[k].off, at + r * d.pack[k].rep_stride, d.pack[k].dtype, p);
      }
  }
}

__global__ void __launch_bounds__(128)
batch_reduce_kernel(const float* __restrict__ src, float* __restrict__ dst, const ast_reduce_desc* __restrict__ descs) {
  const ast_reduce_desc d = descs[blockIdx.y];
  const int c = blockIdx.x * 128 + threadIdx.x;
  if (c >= d.cols) return;
  const float* s = src + d.src_off + c;
  float a0 = 0.f, a1 = 0.f;
  int r = 0;
  for (; r + 1 < d.rows; r += 2) { a0 += s[(long long)r * d.row_stride]; a1 += s[(long long)(r + 1) * d.row_stride]; }
  if (r < d.rows) a0 += s[(long long)r * d.row_stride];
  dst[d.dst_off + c] = a0 + a1;
}

// out[c] += sum over n, h, w of x[n, h, w, c]; one block per (chunk, c, n), rows of the plane split over the chunks
__global__ void __launch_bounds__(256) channel_sum_kernel(Img x, float* __restrict__ out, int chunks) {
  const int c = blockIdx.y, n = blockIdx.z;
  const long long hw = (long long)x.h * x.w;
  const long long beg = hw * blockIdx.x / chunks, end = hw * (blockIdx.x + 1) / chunks;
  float part = 0.f;
  for (long long p = beg + threadIdx.x; p < end; p += 256) {
    const int i = (int)(p / x.w), j = (int)(p - (long long)i * x.w);
    part += ld_elem(x, img_off(x, n, i, j, c));
  }
  __shared__ float red[8];
  part = warp_sum(part);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < 8 ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) atomicAdd(out + c, v);
  }
}

}  // namespace ast

using namespace ast;

extern "C" int ast_adam_step(const ast_param_desc* descs, int32_t n_desc, const int32_t* work, int32_t n_work,
                             float* params, const float* grads, float* exp_avg, float* exp_avg_sq, void* pack_arena,
                             const int64_t* tap_table, ast_adam_state* state, int32_t update, void* stream) {
  AST_CHECK_ARG(descs && work && params && tap_table && n_desc > 0, "ast_adam_step: null argument");
  AST_CHECK_ARG(!update || (grads && exp_avg && exp_avg_sq && state), "ast_adam_step: update needs grads, moments and state");
  if (n_work <= 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (update) {
    launch_k(adam_tick_kernel, 1, 32, 0, s, state);
    launch_k(adam_pack_kernel<true>, n_work, OPT_THREADS, 0, s, descs, work, params, grads, exp_avg, exp_avg_sq, pack_arena,
             tap_table, (const ast_adam_state*)state);
    count_launch(2);
  } else {
    launch_k(adam_pack_kernel<false>, n_work, OPT_THREADS, 0, s, descs, work, params, grads, exp_avg, exp_avg_sq, pack_arena,
             tap_table, (const ast_adam_state*)state);
    count_launch();
  }
  count_work(FAM_OPTIM, 0.0, (update ? 28.0 : 8.0) * n_work * OPT_ITEM);
  AST_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int32_t ast_adam_work_item(void) { return OPT_ITEM; }

extern "C" int ast_batch_reduce(const float* src, float* dst, const ast_reduce_desc* descs, int32_t n_desc, int32_t max_cols,
                                void* stream) {
  AST_CHECK_ARG(src && dst && descs, "ast_batch_reduce: null argument");
  if (n_desc <= 0 || max_cols <= 0) return 0;
  launch_k(batch_reduce_kernel, dim3((max_cols + 127) / 128, n_desc), 128, 0, (cudaStream_t)stream, src, dst, descs);
  count_launch();
  count_work(FAM_OPTIM, 0.0, 0.0);
  AST_CUDA_LAUNCH_CHECK();
  return 0;
}

extern "C" int ast_channel_sum(const ast_image* x, float* out, void* stream) {
  AST_CHECK_ARG(x && out, "ast_channel_sum: null argument");
  AST_CHECK_ARG(x->dtype == AST_F32 || x->dtype == AST_BF16, "ast_channel_sum: fp32 / bf16 images only");
  if (x->n == 0 || x->c == 0 || x->h == 0 || x->w == 0) return 0;
  AST_CHECK_ARG(x->c <= 65535 && x->n <= 65535, "ast_channel_sum: too many channels / images for one grid");
  const long long hw = (long long)x->h * x->w;
  int chunks = (int)((hw + 8191) / 8192);
  if (chunks < 1) chunks = 1;
  launch_k(channel_sum_kernel, dim3(chunks, x->c, x->n), 256, 0, (cudaStream_t)stream, to_img(x), out, chunks);
  count_launch();
  count_work(FAM_POINTWISE, 0.0, img_bytes(x));
  AST_CUDA_LAUNCH_CHECK();
  return 0;
}
