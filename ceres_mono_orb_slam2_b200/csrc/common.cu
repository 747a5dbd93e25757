// Error text and version of the C-ABI library.
#include <cstring>

#include "cmos_common.h"

namespace cmos {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace cmos

extern "C" {
const char* cmos_last_error(void) { return cmos::g_err; }
const char* cmos_version(void) { return "cmos_b200 0.1 sm_100a"; }
}
