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

extern "C" const char* ast_last_error(void) { return ast::g_err; }
extern "C" int ast_abi_version(void) { return AST_ABI_VERSION; }
extern "C" int64_t ast_launch_count(void) { return ast::g_launches.load(); }

namespace ast { int tc_capabilities(); }
extern "C" int ast_capabilities(void) { return ast::tc_capabilities(); }
