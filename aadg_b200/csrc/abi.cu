// Error channel and version of the C ABI (include/aadg_b200.h).
#include "common.cuh"

namespace aadg {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace aadg

extern "C" {
int aadg_version(void) { return AADG_ABI_VERSION; }
const char* aadg_last_error(void) { return aadg::g_err; }
}
