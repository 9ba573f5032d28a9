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

// 32-byte record in mapped, portable host memory that device code fills in before a trap (b2t_trap_record)
static unsigned* g_trap_rec = nullptr;
unsigned* b2t_trap_rec() {
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    if (cudaHostAlloc(&p, 32, cudaHostAllocMapped | cudaHostAllocPortable) == cudaSuccess) {
      memset(p, 0, 32);
      g_trap_rec = (unsigned*)p;
    } else {
      (void)cudaGetLastError();
    }
  }
  return g_trap_rec;
}
static void b2t_trap_text(char* buf, size_t n) {
  buf[0] = 0;
  if (g_trap_rec && g_trap_rec[0])
    snprintf(buf, n, "; device trap record: site 0x%x a=%u b=%u block=(%u,%u) thread=%u", g_trap_rec[0], g_trap_rec[1], g_trap_rec[2],
             g_trap_rec[3], g_trap_rec[4], g_trap_rec[5]);
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
    char rec[160];
    b2t_trap_text(rec, sizeof(rec));
    b2t_set_error("device fault in the kernel launched at %s:%d: %s%s", file, line, cudaGetErrorString(e), rec);
    fprintf(stderr, "b200tok: device fault in the kernel launched at %s:%d: %s%s\n", file, line, cudaGetErrorString(e), rec);
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
__global__ void test_trap_kernel(unsigned* rec) {
  if (rec) b2t_trap_record(rec, 0x7e57u, 0xabcdu, 7u);
  b2t_trap_report("test trap requested through b2t_set_option", 0xabcdu, 7u);
}
int b2t_test_trap(int light) {
  test_trap_kernel<<<1, 1>>>(light ? b2t_trap_rec() : nullptr);
  B2T_LAUNCH_CHECK();
  cudaError_t e = cudaDeviceSynchronize();
  char rec[160];
  b2t_trap_text(rec, sizeof(rec));
  b2t_set_error("test trap: cudaDeviceSynchronize -> %s%s", cudaGetErrorString(e), rec);
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

int b2t_last_device_trap(unsigned* rec6) {
  if (!g_trap_rec || g_trap_rec[0] == 0) return 0;
  if (rec6) memcpy(rec6, g_trap_rec, 6 * sizeof(unsigned));
  return 1;
}

}  // extern "C"
