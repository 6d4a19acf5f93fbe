// Micro-benchmark (scratch, not product): issue rate of tcgen05.mma with both operands resident in shared memory
// (K-major, 128B swizzle), per N and kind, with and without concurrent bulk copies into shared memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I artist_style_transfer_b200/csrc -I include -o scratch/mma_bench scratch/mma_bench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "tc_common.cuh"

using namespace ast;

// dummy definitions the header expects from the library
namespace ast { void set_error(const char*, ...) {} }

struct P {
  int n, iters, kind, copy_kb;   // copy_kb: KB of copies issued per group of 4 MMAs (0 = none)
  int mode;                      // 1 bulk 1-D per-CTA source, 2 bulk 1-D same source for all CTAs, 3 tiled TMA (128 rows x 128 B,
                                 // 256 B pitch) per-CTA source, 4 tiled TMA same source
  int depth;                     // copy groups in flight (<= 8)
  int lanes;                     // 1: the producers are lanes 0..nprod-1 of ONE warp instead of lane 0 of nprod warps
  int commit_every;              // tcgen05.commit after every commit_every groups of 4 MMAs (0 = never)
  int alt_acc, alt_ops;          // alternate the accumulator / the A operand address between groups
  int rand;                      // random operand bits instead of zeros
  int nprod;                     // producer warps (1..4), each issuing iters/nprod groups
  unsigned idesc;
  const char* src;
};

template <int KIND>
__global__ void __launch_bounds__(160, 1) mma_loop(const __grid_constant__ CUtensorMap tm, P p, long long* cycles) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long done_bar, copy_bar[4][8];
  __shared__ unsigned tmem_slot;
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (16384 + 32768 + 4 * 32768) / 16; i += blockDim.x) {
    // p.rand: pseudo-random finite bf16 / tf32 bit patterns (small magnitudes) instead of zeros
    unsigned h = (unsigned)i * 2654435761u + blockIdx.x * 97u;
    unsigned a0 = p.rand ? (0x3c003c00u ^ ((h * 1664525u + 1013904223u) & 0x83ff83ffu)) : 0u;
    unsigned a1 = p.rand ? (0x3c003c00u ^ ((h * 22695477u + 1u) & 0x83ff83ffu)) : 0u;
    ((uint4*)smem)[i] = make_uint4(a0, a1, a0 ^ 0x00110011u, a1 ^ 0x01010101u);
  }
  if (threadIdx.x == 0) {
    mbar_init(&done_bar, 1);
    for (int w = 0; w < 4; ++w) for (int i = 0; i < 8; ++i) mbar_init(&copy_bar[w][i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem = tmem_slot;
  const unsigned a_addr = smem_u32(smem);              // 128 rows x 128 B
  const unsigned b_addr = a_addr + 16384;              // up to 256 rows x 128 B
  if (warp == 0 && lane == 0 && p.n > 0) {
    const unsigned desc_hi = (unsigned)((1024u >> 4) | (1u << 14) | (2u << 29));
    const unsigned a_lo = ((a_addr & 0x3FFFFu) >> 4) | (1u << 16), b_lo = ((b_addr & 0x3FFFFu) >> 4) | (1u << 16);
    const long long t0 = clock64();
    for (int it = 0; it < p.iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        tc_mma<KIND>(tmem + (p.alt_acc ? (unsigned)((it & 1) * 256) : 0u), pack_desc64(a_lo + 2 * k + (p.alt_ops ? (unsigned)((it & 3) * 64) : 0u), desc_hi),
                     pack_desc64(b_lo + 2 * k, desc_hi), p.idesc, (it | k) ? 1u : 0u);
      if (p.commit_every && (it % p.commit_every) == p.commit_every - 1) tc_commit(&copy_bar[3][it & 7]);   // nobody waits on these
    }
    tc_commit(&done_bar);
    mbar_wait(&done_bar, 0);
    cycles[blockIdx.x] = clock64() - t0;
  } else if (p.copy_kb && ((p.lanes == 0 && warp >= 1 && warp <= p.nprod && lane == 0) || (p.lanes == 1 && warp == 1 && lane < p.nprod))) {
    const int w = p.lanes ? lane : warp - 1;
    const unsigned c_addr = b_addr + 32768 + (unsigned)w * 32768u;   // 32 KB landing zone per producer
    const size_t cta_off = (p.mode == 1 || p.mode == 3) ? (size_t)blockIdx.x : 0;
    const int my_iters = p.iters / p.nprod;
    const long long t0 = clock64();
    for (int it = 0; it < my_iters; ++it) {
      const int slot = it % p.depth;
      unsigned long long* cb = &copy_bar[w][slot];
      if (it >= p.depth) mbar_wait(cb, (unsigned)(it / p.depth - 1) & 1u);
      mbar_expect_tx(cb, (unsigned)p.copy_kb * 1024u);
      for (int kb = 0; kb < p.copy_kb; kb += 16) {
        const unsigned dst = c_addr + (unsigned)(kb % 32) * 1024u;
        if (p.mode <= 2)
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(dst), "l"(p.src + (size_t)((it * 64 + kb + w * 16) % 256) * 1024 + cta_off * (256 * 1024)),
                         "r"(16384u), "r"(smem_u32(cb)) : "memory");
        else
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                       ::"r"(dst), "l"(&tm), "r"(smem_u32(cb)), "r"(0), "r"((int)(cta_off * 1024 + ((it * 4 + kb / 16 + w) % 7) * 128)) : "memory");
      }
    }
    for (int d = 0; d < p.depth && d < my_iters; ++d) {      // drain
      const int it = my_iters - 1 - d;
      mbar_wait(&copy_bar[w][it % p.depth], (unsigned)(it / p.depth) & 1u);
    }
    cycles[gridDim.x * (1 + w) + blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

int main() {
  int dev = 0, sms = 0, khz = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  char* src;
  cudaMalloc(&src, ((size_t)sms + 2) * 512 * 1024);
  cudaMemset(src, 0, ((size_t)sms + 2) * 512 * 1024);
  long long* cyc;
  cudaMallocManaged(&cyc, sizeof(long long) * sms * 5);
  const size_t smem = 16384 + 32768 + 4 * 32768 + 1024;
  // tiled source: rows of 64 bf16 (128 B) at a 256 B pitch, 1024 rows (256 KB) per CTA
  alignas(64) CUtensorMap tm;
  {
    EncodeTiledFn encode = get_encode();
    cuuint64_t dims[2] = {64, (cuuint64_t)sms * 1024 + 1024};
    cuuint64_t strides[1] = {256};
    cuuint32_t box[2] = {64, 128};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, src, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  }
  cudaFuncSetAttribute(mma_loop<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(mma_loop<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  printf("SMs %d, clock %d kHz\n", sms, khz);
  printf("%5s %5s %5s %6s %8s %12s %12s %10s %8s\n", "kind", "N", "mode", "depth", "copyKB/4", "clk/MMA", "TFLOP/s", "copy TB/s", "B/clk/SM");
  struct Cfg { int kind, n, mode, depth, copy_kb, nprod, lanes, commit_every, alt_acc, alt_ops, m, rand; };
  std::vector<Cfg> cfgs;
  for (int rnd : {0, 1})
    for (int kind : {0, 1})
      for (int n : {128, 256}) cfgs.push_back({kind, n, 0, 8, 0, 1, 0, 1, 1, 1, 128, rnd});
  for (int rnd : {0, 1}) cfgs.push_back({0, 256, 3, 8, 32, 3, 1, 1, 1, 1, 128, rnd});
  for (const Cfg& c : cfgs) {
      {
        const int kind = c.kind, n = c.n, copy_kb = c.copy_kb;
        P p;
        p.mode = c.mode; p.depth = c.depth; p.nprod = c.nprod; p.lanes = c.lanes; p.commit_every = c.commit_every; p.rand = c.rand; p.alt_acc = c.alt_acc; p.alt_ops = c.alt_ops;
        p.n = n; p.iters = 20000; p.kind = kind; p.copy_kb = copy_kb; p.src = src;
        const unsigned fmt = kind ? 2u : 1u;
        p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((unsigned)(n >> 3) << 17) | (((unsigned)c.m >> 4) << 24);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int rep = 0; rep < 2; ++rep) {
          cudaEventRecord(e0);
          if (kind == 0) mma_loop<0><<<sms, 160, smem>>>(tm, p, cyc); else mma_loop<1><<<sms, 160, smem>>>(tm, p, cyc);
          cudaEventRecord(e1);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
        }
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        double avg = 0, cavg = 0;
        for (int i = 0; i < sms; ++i) { avg += (double)cyc[i]; cavg += (double)cyc[sms + i]; }
        avg /= sms; cavg /= sms;
        const double mmas = 4.0 * p.iters;
        const double groups = (double)(p.iters / p.nprod);
        printf("rand %d %5s M=%3d N=%3d commit/%d altacc %d altops %d mode %d lanes %d nprod %d copy %2d KB/group | MMA %6.1f clk/MMA | producer %7.1f clk/group -> %6.1f B/clk/SM | kernel %.3f ms\n",
               c.rand, kind ? "tf32" : "bf16", c.m, n, c.commit_every, c.alt_acc, c.alt_ops, c.mode, c.lanes, c.nprod, copy_kb, n ? avg / mmas : 0.0, copy_kb ? cavg / groups : 0.0,
               copy_kb ? copy_kb * 1024.0 * p.nprod / (cavg / groups) : 0.0, ms);
      }
  }
  return 0;
}
