// api.cu -- library bookkeeping: error text, launch counter, device properties.
#include <atomic>
#include <cstdarg>
#include <cstring>
#include "common.cuh"

namespace mvd {
static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};
static std::atomic<unsigned long long> g_fallbacks{0};
static int g_num_sms = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_fallback() { g_fallbacks.fetch_add(1ull, std::memory_order_relaxed); }
void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }
int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      g_num_sms = n;
    else
      g_num_sms = 148;
  }
  return g_num_sms;
}
}  // namespace mvd

// busy-wait of `cycles` SM clocks on one thread: measurement aid (bench.py queues it in front of an event-timed launch so
// that the host has enqueued the whole bracket before the GPU reaches it -- otherwise the events also measure how long
// the GPU waited for the Python / ctypes launch path)
__global__ void spin_kernel(long long cycles) {
  const long long t0 = clock64();
  while (clock64() - t0 < cycles) {
  }
}

extern "C" {
int mvd_spin(long long cycles, mvd_stream_t stream) {
  if (cycles <= 0) return MVD_OK;
  spin_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(cycles);
  return cudaGetLastError() == cudaSuccess ? MVD_OK : MVD_ERR_CUDA;
}
int mvd_version(void) { return 100; }
const char* mvd_last_error(void) { return mvd::g_err; }
unsigned long long mvd_launch_count(void) { return mvd::g_launches.load(); }
void mvd_reset_launch_count(void) { mvd::g_launches.store(0); }
unsigned long long mvd_fallback_count(void) { return mvd::g_fallbacks.load(); }
void mvd_reset_fallback_count(void) { mvd::g_fallbacks.store(0); }
int mvd_shutdown(void) { return MVD_OK; }
}
