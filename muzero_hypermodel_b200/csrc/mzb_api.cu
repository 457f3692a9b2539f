// Library-wide plumbing of libmzb200: version, thread-local error text, launch counter, host Philox.
#include <stdarg.h>

#include <atomic>

#include "mzb_common.cuh"

static thread_local char g_error[512] = "";
static std::atomic<uint64_t> g_launches{0};

void mzb_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

void mzb_count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

extern "C" {

int mzb_version(void) { return MZB_VERSION; }
const char* mzb_last_error(void) { return g_error; }
uint64_t mzb_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
void mzb_reset_launch_count(void) { g_launches.store(0, std::memory_order_relaxed); }

void mzb_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
  const Philox4 r = philox4x32_10(c0, c1, c2, c3, k0, k1);
  out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

}  // extern "C"
