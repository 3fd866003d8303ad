// libmis_b200.so -- version / error plumbing of the C ABI (include/mis_b200.h).
#include "common.cuh"

namespace mis {

char* last_error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buffer(), 512, fmt, ap);
  va_end(ap);
  return code;
}

}  // namespace mis

extern "C" int mis_version(void) { return MIS_ABI_VERSION; }

extern "C" const char* mis_last_error(void) { return mis::last_error_buffer(); }
