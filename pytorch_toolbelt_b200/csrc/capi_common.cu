// Error plumbing and version for the tb200 C-ABI library.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace tb {

char* err_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

int cuda_fail(cudaError_t e, const char* what) {
  snprintf(err_buf(), 512, "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
  return (int)e;
}

}  // namespace tb

extern "C" {
int tb_version(void) { return TB200_VERSION; }
const char* tb_last_error(void) { return tb::err_buf(); }
}
