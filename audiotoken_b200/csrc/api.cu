// Library-level entry points: version, error text, device check, launch counter.
#include <stdarg.h>
#include <string.h>
#include "common.cuh"

static thread_local char g_err[512] = "";
static thread_local int g_launches = 0;

void b2t_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void b2t_count_launch(int n) { g_launches += n; }
void b2t_reset_launch_count() { g_launches = 0; }

int b2t_num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

int b2t_arch_ok() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    b2t_set_error("no CUDA device: b200tok has no CPU fallback");
    return B2T_ERR_CUDA;
  }
  return b2t_device_check(dev);
}

extern "C" {

int b2t_version(void) { return 100; }

const char* b2t_last_error(void) { return g_err; }

int b2t_device_check(int device) {
  int major = 0, minor = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device) != cudaSuccess ||
      cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device) != cudaSuccess) {
    b2t_set_error("cannot query device %d: b200tok has no CPU fallback", device);
    return B2T_ERR_CUDA;
  }
  if (major != 10) {
    b2t_set_error("device %d is sm_%d%d; b200tok is built for sm_100a only", device, major, minor);
    return B2T_ERR_ARCH;
  }
  return B2T_OK;
}

int b2t_last_launch_count(void) { return g_launches; }

}  // extern "C"
