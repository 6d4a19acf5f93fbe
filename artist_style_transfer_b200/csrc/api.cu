// Error plumbing and process-wide counters of libast_b200.so.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace ast {
static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace ast

namespace ast {
static std::atomic<long long> g_fam_launches[FAM_COUNT];
static std::atomic<double> g_fam_flops[FAM_COUNT], g_fam_bytes[FAM_COUNT];
static void atomic_add(std::atomic<double>& a, double v) {
  double cur = a.load(std::memory_order_relaxed);
  while (!a.compare_exchange_weak(cur, cur + v, std::memory_order_relaxed)) {}
}
void count_work(int family, double flops, double bytes) {
  if (family < 0 || family >= FAM_COUNT) return;
  g_fam_launches[family].fetch_add(1, std::memory_order_relaxed);
  atomic_add(g_fam_flops[family], flops);
  atomic_add(g_fam_bytes[family], bytes);
}
static const char* const kFamilyNames[FAM_COUNT] = {
  "conv_tc", "conv_px", "conv_ws", "conv_hx", "conv_simt", "wgrad_tc", "wgrad_thin", "wgrad_simt", "gram_tc", "gram_simt",
  "in_apply", "in_bwd", "in_stats", "pool", "mse", "pointwise", "optim", "conv_st"};
}  // namespace ast

extern "C" int ast_family_count(void) { return ast::FAM_COUNT; }
extern "C" const char* ast_family_name(int family) {
  return (family >= 0 && family < ast::FAM_COUNT) ? ast::kFamilyNames[family] : "";
}
extern "C" int ast_family_stats(int family, int64_t* launches, double* flops, double* bytes) {
  if (family < 0 || family >= ast::FAM_COUNT) { ast::set_error("ast_family_stats: family %d out of range", family); return -1; }
  if (launches) *launches = ast::g_fam_launches[family].load();
  if (flops) *flops = ast::g_fam_flops[family].load();
  if (bytes) *bytes = ast::g_fam_bytes[family].load();
  return 0;
}
extern "C" void ast_family_reset(void) {
  for (int f = 0; f < ast::FAM_COUNT; ++f) { ast::g_fam_launches[f] = 0; ast::g_fam_flops[f] = 0.0; ast::g_fam_bytes[f] = 0.0; }
}

extern "C" const char* ast_last_error(void) { return ast::g_err; }
extern "C" int ast_abi_version(void) { return AST_ABI_VERSION; }
extern "C" int64_t ast_launch_count(void) { return ast::g_launches.load(); }

namespace ast { int tc_capabilities(); }
extern "C" int ast_capabilities(void) { return ast::tc_capabilities(); }
