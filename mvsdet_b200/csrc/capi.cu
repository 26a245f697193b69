// Library-level pieces of the C ABI (include/mvsdet_b200.h): status strings,
// thread-local error text, tuning knobs, launch counter.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#define MVSD_STR2(x) #x
#define MVSD_STR(x) MVSD_STR2(x)
#include "common.cuh"

namespace mvsd {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};
static std::atomic<int> g_tuning[8];

int fail(int status, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return status;
}

// Launch-configuration errors of OUR launch are non-sticky: they are reported through the status
// and cleared, so they do not resurface in the caller's next CUDA call.  A sticky error (an
// asynchronous fault of an earlier kernel, possibly the caller's own) cannot be cleared by
// cudaGetLastError and stays visible to the caller; the message says which case it is.
int check_launch(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e == cudaSuccess) return MVSD_OK;
  cudaGetLastError();
  const bool sticky = cudaPeekAtLastError() != cudaSuccess;
  return fail(MVSD_ERR_CUDA, "%s: %s (%s)%s", what, cudaGetErrorName(e), cudaGetErrorString(e),
              sticky ? " -- sticky context error, raised by an earlier kernel (not necessarily one of "
                       "this library's)" : "");
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int tuning(int key) { return (key >= 0 && key < 8) ? g_tuning[key].load() : 0; }

}  // namespace mvsd

extern "C" int mvsd_abi_version(void) { return MVSD_ABI_VERSION; }

extern "C" const char* mvsd_build_info(void) {
  return "mvsdet_b200 sm_100a, nvcc " MVSD_STR(__CUDACC_VER_MAJOR__) "." MVSD_STR(__CUDACC_VER_MINOR__)
         ", built " __DATE__;
}

extern "C" const char* mvsd_status_string(int status) {
  switch (status) {
    case MVSD_OK: return "ok";
    case MVSD_ERR_INVALID_ARG: return "invalid argument";
    case MVSD_ERR_UNSUPPORTED: return "unsupported configuration";
    case MVSD_ERR_CUDA: return "CUDA error";
    default: return "unknown status";
  }
}

extern "C" const char* mvsd_last_error(void) { return mvsd::g_err; }

extern "C" int mvsd_set_tuning(int key, int value) {
  if (key < 0 || key >= 8) return -1;
  return mvsd::g_tuning[key].exchange(value);
}

extern "C" int64_t mvsd_launch_count(void) { return mvsd::g_launches.load(); }
