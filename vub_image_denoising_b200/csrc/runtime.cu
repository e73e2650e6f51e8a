// runtime.cu — error reporting and device queries of libb200dn.
#include "common.cuh"

#include <mutex>
#include <string.h>
#include <stdlib.h>

namespace b200dn {

namespace {
thread_local char g_err[512] = "";
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", static_cast<int>(e), cudaGetErrorString(e), what);
  return B200DN_E_CUDA;
}

namespace {
constexpr int kMaxDev = kMaxDevices;
int g_sm_count[kMaxDev];
int g_cc_major[kMaxDev];
bool g_have[kMaxDev];

int load_props(int* dev_out) {
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
  if (dev < 0 || dev >= kMaxDev) {
    set_error("device ordinal %d out of range", dev);
    return B200DN_E_CUDA;
  }
  if (!g_have[dev]) {
    int sms = 0, major = 0;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceGetAttribute(SM count)");
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceGetAttribute(cc major)");
    g_sm_count[dev] = sms;
    g_cc_major[dev] = major;
    g_have[dev] = true;
  }
  *dev_out = dev;
  return 0;
}
}  // namespace

int device_sm_count() {
  int dev = 0;
  if (int rc = load_props(&dev)) return rc;
  return g_sm_count[dev];
}

int require_sm100() {
  int dev = 0;
  if (int rc = load_props(&dev)) return rc;
  if (g_cc_major[dev] != 10) {
    set_error("libb200dn needs a compute-capability 10.x (Blackwell B200) device, found %d.x — there is no fallback path",
              g_cc_major[dev]);
    return B200DN_E_CUDA;
  }
  return 0;
}

int ensure_max_dyn_smem(SmemOptIn& st, const void* const* kernels, int n, int bytes, const char* what) {
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
  if (dev < 0 || dev >= kMaxDevices) {
    set_error("device ordinal %d out of range", dev);
    return B200DN_E_CUDA;
  }
  if (st.done[dev]) return 0;
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  if (st.done[dev]) return 0;
  for (int i = 0; i < n; ++i) {
    e = cudaFuncSetAttribute(kernels[i], cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return cuda_fail(e, what);
  }
  st.done[dev] = true;
  return 0;
}

cudaError_t launch_pdl(const void* kernel, int grid, int threads, size_t smem, cudaStream_t stream, void* params,
                       int cluster, bool cooperative) {
  static const int pdl = [] {
    const char* e = getenv("B200DN_PDL");
    return e ? atoi(e) : 1;
  }();
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(static_cast<unsigned>(threads));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[3];
  int n = 0;
  if (cooperative) {
    attr[n].id = cudaLaunchAttributeCooperative;
    attr[n].val.cooperative = 1;
    ++n;
  }
  if (cluster > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = static_cast<unsigned>(cluster);
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  void* args[1] = {params};
  return cudaLaunchKernelExC(&cfg, kernel, args);
}

}  // namespace b200dn

extern "C" const char* b200dn_last_error(void) { return b200dn::g_err; }
extern "C" int b200dn_abi_version(void) { return B200DN_ABI_VERSION; }
extern "C" int b200dn_sm_count(void) { return b200dn::device_sm_count(); }
