// C-ABI plumbing: thread-local error string, version.
#include "common.cuh"
#include <stdarg.h>

static thread_local char g_err[512] = "";

void lcgan_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* lcgan_last_error(void) { return g_err; }
extern "C" int lcgan_version(void) { return 2; }

// Deterministic mode: ordered (turn-taking) reductions instead of fp32 atomics - see common.cuh.
static int g_det = 0;
bool lcgan_det_enabled() { return g_det != 0; }
extern "C" int lcgan_set_deterministic(int on) {
  const int old = g_det;
  g_det = on ? 1 : 0;
  return old;
}
