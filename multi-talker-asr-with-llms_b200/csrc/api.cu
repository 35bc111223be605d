// Library-wide plumbing of the C ABI: version, thread-local error string, device properties.
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "common.cuh"

namespace mtasr {

static thread_local char g_err[1024] = "";

char* last_error_buf() { return g_err; }

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

static std::atomic<int> g_sm_budget{0};

static int sm_count() {
  static int n = 0;
  static std::once_flag once;
  std::call_once(once, [] {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  });
  return n;
}

int num_sms() {
  const int n = sm_count();
  const int b = g_sm_budget.load(std::memory_order_relaxed);
  return (b > 0 && b < n) ? b : n;
}

// Schedulable units of a persistent kernel: CTAs (ncta = 1) or CTA pairs (ncta = 2, cluster of two SMs of one TPC).
// With R SMs left to a foreign kernel (NCCL), every one of its CTAs may sit on a different TPC and break a pair, so only
// (total - 2 R) / 2 pairs are guaranteed to be placeable at once; a pair that is not waits for the foreign kernel and the
// statically scheduled tiles it owns finish late (measured: 100 vs 116 ms per step, bimodal, with (total - R) / 2 pairs).
int num_units(int ncta) {
  const int usable = num_sms();
  if (ncta <= 1) return usable;
  const int total = sm_count();
  const int pairs = (total - 2 * (total - usable)) / 2;
  return pairs > 1 ? pairs : 1;
}

bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("MTASR_PDL");
    return e == nullptr || e[0] != '0';
  }();
  return on;
}

}  // namespace mtasr

extern "C" int mtasr_version(void) { return 100; }
extern "C" int mtasr_set_sm_budget(int32_t n_sms) {
  if (n_sms < 0) return mtasr::set_error(MTASR_ERR_INVALID_ARG, "set_sm_budget: negative SM count");
  mtasr::g_sm_budget.store(n_sms & ~1);   // even: CTA pairs
  return mtasr::num_sms();
}
extern "C" const char* mtasr_last_error_string(void) { return mtasr::last_error_buf(); }
