// Micro-benchmark (scratch, not product): cost of tcgen05.mma (cta_group::1, both operands in shared memory, K-major) when
// EVERY instruction reads different operand tiles - the situation of the real conv kernels - as a function of M, N, the
// swizzle / row width (128-byte rows vs the 64-byte rows of 32-channel bf16 tensors) and random vs zero data.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I artist_style_transfer_b200/csrc -I include -o scratch/mma_bench2 scratch/mma_bench2.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "tc_common.cuh"

using namespace ast;
namespace ast { void set_error(const char*, ...) {} }

struct P {
  int m, n, iters, kind, rowb, na, nb, commit_every, rand;
  unsigned idesc;
};

// smem: NA A-tiles of 128 rows x rowb bytes, then NB B-tiles of 256 rows x rowb bytes
template <int KIND>
__global__ void __launch_bounds__(128, 1) mma_loop(P p, long long* cycles) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long done_bar, junk_bar[8];
  __shared__ unsigned tmem_slot;
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int a_tile = 128 * p.rowb, b_tile = 256 * p.rowb;
  const int total16 = (p.na * a_tile + p.nb * b_tile) / 16;
  for (int i = threadIdx.x; i < total16; i += blockDim.x) {
    unsigned h = (unsigned)i * 2654435761u + blockIdx.x * 97u;
    unsigned a0 = p.rand ? (0x3c003c00u ^ ((h * 1664525u + 1013904223u) & 0x83ff83ffu)) : 0u;
    unsigned a1 = p.rand ? (0x3c003c00u ^ ((h * 22695477u + 1u) & 0x83ff83ffu)) : 0u;
    ((uint4*)smem)[i] = make_uint4(a0, a1, a0 ^ 0x00110011u, a1 ^ 0x01010101u);
  }
  if (threadIdx.x == 0) {
    mbar_init(&done_bar, 1);
    for (int i = 0; i < 8; ++i) mbar_init(&junk_bar[i], 1);
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
  const unsigned a0 = smem_u32(smem), b0 = a0 + p.na * a_tile;
  if (warp == 0 && lane == 0) {
    const unsigned layout = p.rowb == 128 ? 2u : 4u;
    const unsigned desc_hi = (unsigned)(((8u * p.rowb) >> 4) | (1u << 14) | (layout << 29));
    const int kmma = p.rowb / 32;
    const unsigned a_base = ((a0 & 0x3FFFFu) >> 4) | (1u << 16), b_base = ((b0 & 0x3FFFFu) >> 4) | (1u << 16);
    const unsigned a_step = (unsigned)a_tile >> 4, b_step = (unsigned)b_tile >> 4;
    const unsigned na_mask = (unsigned)p.na - 1, nb_mask = (unsigned)p.nb - 1;     // powers of two
    const unsigned ce_mask = p.commit_every ? (unsigned)p.commit_every - 1 : 0xffffffffu;
    const long long t0 = clock64();
    for (int it = 0; it < p.iters; ++it) {
      const unsigned a_lo = a_base + ((unsigned)it & na_mask) * a_step;
      const unsigned b_lo = b_base + ((unsigned)it & nb_mask) * b_step;
      const unsigned d = tmem + (unsigned)((it & 1) * 256);
      tc_mma<KIND>(d, pack_desc64(a_lo, desc_hi), pack_desc64(b_lo, desc_hi), p.idesc, it > 1 ? 1u : 0u);
      tc_mma<KIND>(d, pack_desc64(a_lo + 2, desc_hi), pack_desc64(b_lo + 2, desc_hi), p.idesc, 1u);
      if (kmma == 4) {
        tc_mma<KIND>(d, pack_desc64(a_lo + 4, desc_hi), pack_desc64(b_lo + 4, desc_hi), p.idesc, 1u);
        tc_mma<KIND>(d, pack_desc64(a_lo + 6, desc_hi), pack_desc64(b_lo + 6, desc_hi), p.idesc, 1u);
      }
      if (p.commit_every && ((unsigned)it & ce_mask) == ce_mask) tc_commit(&junk_bar[it & 7]);
    }
    const long long t1 = clock64();
    tc_commit(&done_bar);
    mbar_wait(&done_bar, 0);
    cycles[blockIdx.x] = clock64() - t0;
    cycles[gridDim.x + blockIdx.x] = t1 - t0;
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
  long long* cyc;
  cudaMallocManaged(&cyc, sizeof(long long) * sms * 2);
  printf("SMs %d, clock %d kHz: cycles per tcgen05.mma (32 bytes of K), operands cycling over NA A-tiles / NB B-tiles\n", sms, khz);
  struct Cfg { int kind, m, n, rowb, na, nb, ce, rnd; };
  std::vector<Cfg> cfgs;
  for (int rowb : {128, 64})
    for (int m : {64, 128})
      for (int n : {32, 64, 128, 256}) {
        cfgs.push_back({0, m, n, rowb, 1, 1, 4, 1});      // same operands every time (what r01's mma_bench measured)
        cfgs.push_back({0, m, n, rowb, 4, 2, 4, 1});      // different operands every group
      }
  for (int ce : {0, 1, 2, 8}) cfgs.push_back({0, 128, 256, 128, 4, 2, ce, 1});
  for (int ce : {0, 1, 2, 8}) cfgs.push_back({0, 128, 64, 128, 4, 2, ce, 1});
  cfgs.push_back({1, 128, 256, 128, 4, 2, 4, 1});
  cfgs.push_back({1, 128, 64, 128, 4, 2, 4, 1});
  cfgs.push_back({0, 128, 256, 128, 4, 2, 4, 0});
  cfgs.push_back({0, 128, 256, 128, 4, 1, 4, 1});             // only A changes
  cfgs.push_back({0, 128, 256, 128, 1, 2, 4, 1});             // only B changes
  for (const Cfg& c : cfgs) {
    P p;
    p.m = c.m; p.n = c.n; p.iters = 20000; p.kind = c.kind; p.rowb = c.rowb; p.na = c.na; p.nb = c.nb; p.commit_every = c.ce; p.rand = c.rnd;
    const unsigned fmt = c.kind ? 2u : 1u;
    p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((unsigned)(c.n >> 3) << 17) | (((unsigned)c.m >> 4) << 24);
    const size_t smem = (size_t)c.na * 128 * c.rowb + (size_t)c.nb * 256 * c.rowb + 1024;
    cudaFuncSetAttribute(mma_loop<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(mma_loop<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int rep = 0; rep < 2; ++rep) {
      if (c.kind == 0) mma_loop<0><<<sms, 128, smem>>>(p, cyc); else mma_loop<1><<<sms, 128, smem>>>(p, cyc);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    }
    double avg = 0, iavg = 0;
    for (int i = 0; i < sms; ++i) { avg += (double)cyc[i]; iavg += (double)cyc[sms + i]; }
    avg /= sms; iavg /= sms;
    const double mmas = (double)p.iters * (c.rowb / 32);
    const double clk = avg / mmas;
    printf("%s M=%3d N=%3d rows %3d B  NA %d NB %d commit/%d %s | %6.1f clk/MMA (issue loop alone %6.1f) | (M+N)*32B/clk = %5.1f B/clk | %6.0f TFLOP/s at this rate\n", c.kind ? "tf32" : "bf16",
           c.m, c.n, c.rowb, c.na, c.nb, c.ce, c.rnd ? "rand" : "zero", clk, iavg / mmas, (c.m + c.n) * 32.0 / clk,
           2.0 * c.m * c.n * (c.kind ? 8 : 16) / clk * sms * (khz * 1e3) / 1e12);
  }
  return 0;
}
