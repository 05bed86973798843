// libbtb200: error plumbing, version and device checks of the C ABI (include/btb200.h).
#include "bt_common.cuh"
#include <string.h>

static thread_local char g_err[512] = "";

void bt_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static thread_local int g_ptr_checks = 1;

int bt_check_device_ptr(const void* p, const char* name) {
  if (p == nullptr) {
    bt_set_error("%s is NULL", name);
    return BT_ERR_BAD_POINTER;
  }
  if (!g_ptr_checks) return BT_OK;   // bt_set_pointer_checks(0): the caller vouches (CUDA-graph capture)
  cudaPointerAttributes a;
  cudaError_t e = cudaPointerGetAttributes(&a, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    bt_set_error("%s: cudaPointerGetAttributes failed: %s", name, cudaGetErrorString(e));
    return BT_ERR_BAD_POINTER;
  }
  if (a.type != cudaMemoryTypeDevice && a.type != cudaMemoryTypeManaged) {
    bt_set_error("%s is not a device pointer (libbtb200 has no CPU path)", name);
    return BT_ERR_BAD_POINTER;
  }
  return BT_OK;
}

extern "C" {

int bt_version(void) { return BT_VERSION; }

int bt_set_pointer_checks(int enabled) {
  const int prev = g_ptr_checks;
  g_ptr_checks = enabled ? 1 : 0;
  return prev;
}

const char* bt_last_error(void) { return g_err; }

int bt_device_check(void) {
  int dev = 0;
  BT_CHECK_CUDA(cudaGetDevice(&dev));
  int major = 0, minor = 0;
  BT_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  BT_CHECK_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  BT_REQUIRE(major == 10, BT_ERR_UNSUPPORTED,
             "libbtb200 is built for sm_100a only; device %d is sm_%d%d", dev, major, minor);
  return BT_OK;
}

int bt_sm_count(void) {
  int dev = 0, n = 0;
  BT_CHECK_CUDA(cudaGetDevice(&dev));
  BT_CHECK_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  return n;
}

}  // extern "C"
