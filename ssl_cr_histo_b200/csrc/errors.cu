#include <stdarg.h>
#include <stdio.h>

#include "launch.h"

namespace b2n {

static thread_local char g_err[512] = "";

int set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return 1;
}
const char* last_error() { return g_err; }

}  // namespace b2n
