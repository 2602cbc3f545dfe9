// Error text plumbing for the C-ABI (thread-local last error).
#include <stdarg.h>

#include "common.cuh"

namespace mtam {
std::string& last_error_slot() {
  static thread_local std::string s;
  return s;
}
long long& launch_counter() {
  static long long n = 0;
  return n;
}
int set_error(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error_slot() = buf;
  return code;
}
}  // namespace mtam

extern "C" long long mtam_launch_count(void) { return mtam::launch_counter(); }
