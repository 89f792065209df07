// Library-level entry points of the C-ABI: version, last-error text.
#include "common.cuh"
#include <string.h>

static thread_local char g_last_error[512] = "";

void tfb_set_last_error(const char* msg) {
  strncpy(g_last_error, msg ? msg : "", sizeof(g_last_error) - 1);
  g_last_error[sizeof(g_last_error) - 1] = 0;
}

TFB_API const char* tfb_last_error(void) { return g_last_error; }
TFB_API int tfb_abi_version(void) { return 1; }
