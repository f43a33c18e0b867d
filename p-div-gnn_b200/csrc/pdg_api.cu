// Library-wide state: last-error string, version, SM count.
#include <mutex>
#include <stdarg.h>

#include "pdg_common.cuh"

namespace pdg {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int num_sms() {
  static int sms[64] = {0};
  static std::mutex mu;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  std::lock_guard<std::mutex> lk(mu);
  if (sms[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    sms[dev] = v;
  }
  return sms[dev];
}
}  // namespace pdg

extern "C" const char* pdg_last_error(void) { return pdg::g_err; }
extern "C" int pdg_version(void) { return 100; }
extern "C" int pdg_num_sms(void) { return pdg::num_sms(); }
