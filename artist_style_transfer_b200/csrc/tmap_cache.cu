// Host-side caches of the tensor-core kernels: encoded TMA descriptors and per-kernel shared-memory limits.
#include <mutex>
#include <string>
#include <unordered_map>

#include "tc_common.cuh"

namespace ast {

namespace {
struct TmKey {
  unsigned long long v[20];
  bool operator==(const TmKey& o) const { return memcmp(v, o.v, sizeof(v)) == 0; }
};
struct TmHash {
  size_t operator()(const TmKey& k) const {
    unsigned long long h = 1469598103934665603ull;
    for (unsigned long long x : k.v) { h ^= x; h *= 1099511628211ull; }
    return (size_t)h;
  }
};
struct alignas(64) TmVal { CUtensorMap m; };
std::mutex g_tm_mu;
std::unordered_map<TmKey, TmVal, TmHash> g_tm;
std::mutex g_smem_mu;
std::unordered_map<unsigned long long, size_t> g_smem;
}  // namespace

int cached_tensor_map(EncodeTiledFn encode, CUtensorMap* out, CUtensorMapDataType dt, int rank, void* ptr,
                      const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box, const cuuint32_t* estr,
                      CUtensorMapSwizzle sw, CUtensorMapL2promotion l2) {
  TmKey k;
  memset(&k, 0, sizeof(k));
  k.v[0] = (unsigned long long)(uintptr_t)ptr;
  k.v[1] = ((unsigned long long)dt << 40) | ((unsigned long long)rank << 32) | ((unsigned long long)sw << 8) | (unsigned long long)l2;
  for (int i = 0; i < rank; ++i) { k.v[2 + i] = dims[i]; k.v[10 + i] = ((unsigned long long)box[i] << 32) | estr[i]; }
  for (int i = 0; i + 1 < rank; ++i) k.v[6 + i] = strides[i];
  {
    std::lock_guard<std::mutex> lk(g_tm_mu);
    auto it = g_tm.find(k);
    if (it != g_tm.end()) { memcpy(out, &it->second.m, sizeof(CUtensorMap)); return 0; }
  }
  alignas(64) CUtensorMap m;
  const CUresult r = encode(&m, dt, (cuuint32_t)rank, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, l2,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed: %d (rank %d, box0 %u)", (int)r, rank, box[0]); return (int)r; }
  {
    std::lock_guard<std::mutex> lk(g_tm_mu);
    if (g_tm.size() > 8192) g_tm.clear();          // bounded: shapes of a training job repeat, this is only a safety valve
    TmVal v; memcpy(&v.m, &m, sizeof(m));
    g_tm.emplace(k, v);
  }
  memcpy(out, &m, sizeof(m));
  return 0;
}

cudaError_t set_max_smem_impl(const void* kernel, size_t smem) {
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long key = (unsigned long long)(uintptr_t)kernel ^ ((unsigned long long)dev << 56);
  std::lock_guard<std::mutex> lk(g_smem_mu);
  auto it = g_smem.find(key);
  if (it != g_smem.end() && it->second >= smem) return cudaSuccess;
  const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) g_smem[key] = smem;
  return e;
}

}  // namespace ast
