// Last-error string + ABI version of libtfpp.so.
#include <stdarg.h>
#include <stdio.h>

#include "../../include/tfpp.h"

static thread_local char g_err[512] = "";

extern "C" void tfpp_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
extern "C" const char* tfpp_last_error(void) { return g_err; }
extern "C" int tfpp_abi_version(void) { return 2; }
