// Library-level entry points: version, error text, device check, launch counter.
#include <stdarg.h>
#include <stdlib.h>
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

// per device: a process may drive several GPUs (AudioToken(device='cuda:0') and ('cuda:1'))
int b2t_device_index() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < B2T_MAX_DEVICES) ? dev : 0;
}

int b2t_num_sms() {
  static int sms[B2T_MAX_DEVICES] = {};
  const int dev = b2t_device_index();
  if (sms[dev] == 0) {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    sms[dev] = v > 0 ? v : 148;
  }
  return sms[dev];
}

static int g_debug_sync = -1;
bool b2t_debug_sync() {
  if (g_debug_sync < 0) {
    const char* e = getenv("B2T_DEBUG_SYNC");
    g_debug_sync = (e && *e && *e != '0') ? 1 : 0;
  }
  return g_debug_sync != 0;
}
void b2t_set_debug_sync(int on) { g_debug_sync = on ? 1 : 0; }
int b2t_debug_sync_check(const char* file, int line) {
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    b2t_set_error("device fault in the kernel launched at %s:%d: %s", file, line, cudaGetErrorString(e));
    fprintf(stderr, "b200tok: device fault in the kernel launched at %s:%d: %s\n", file, line, cudaGetErrorString(e));
    return B2T_ERR_CUDA;
  }
  return B2T_OK;
}

int b2t_arch_ok() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    b2t_set_error("no CUDA device: b200tok has no CPU fallback");
    return B2T_ERR_CUDA;
  }
  return b2t_device_check(dev);
}

// developer hook (b2t_set_option("test_trap", 1)): one thread runs the device trap report, so that the host-side
// reporting of a protocol time-out can be checked on a GPU without breaking a protocol
__global__ void test_trap_kernel() { b2t_trap_report("test trap requested through b2t_set_option", 0xabcdu, 7u); }
int b2t_test_trap() {
  test_trap_kernel<<<1, 1>>>();
  B2T_LAUNCH_CHECK();
  cudaError_t e = cudaDeviceSynchronize();
  b2t_set_error("test trap: cudaDeviceSynchronize -> %s", cudaGetErrorString(e));
  return e == cudaSuccess ? B2T_OK : B2T_ERR_CUDA;
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
