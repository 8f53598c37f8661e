// Error reporting and version entry points of the C-ABI.
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

namespace gnf {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

#ifndef GNF_EMU
bool pdl_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("GNF_PDL"); on = (e && e[0] == '0') ? 0 : 1; }      // GNF_PDL=0: measurement (plain launches)
  return on == 1;
}

const Branches& branches() {
  constexpr int kMaxDev = 16;
  static Branches pool[kMaxDev];
  static bool made[kMaxDev] = {};
  static Branches none = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev) return none;
  if (!made[dev]) {
    Branches& b = pool[dev];
    b.ok = true;
    for (int k = 0; k < Branches::kSide; ++k) {
      b.ok = b.ok && cudaStreamCreateWithFlags(&b.side[k], cudaStreamNonBlocking) == cudaSuccess;
      b.ok = b.ok && cudaEventCreateWithFlags(&b.fork_ev[k], cudaEventDisableTiming) == cudaSuccess;
      b.ok = b.ok && cudaEventCreateWithFlags(&b.join_ev[k], cudaEventDisableTiming) == cudaSuccess;
    }
    if (!b.ok) cudaGetLastError();
    made[dev] = true;
  }
  return pool[dev];
}
#endif

}  // namespace gnf

extern "C" {
int gnf_version(void) { return 100; }
const char* gnf_last_error(void) { return gnf::g_err; }
int gnf_has_device_code(void) {
#ifdef GNF_EMU
  return 0;
#else
  return 1;
#endif
}
}
