// Library-wide state of libfgs_b200.so: error string, launch counter, device selection.
#include "common.cuh"

namespace fgs {
thread_local char g_err[256] = "";
int64_t g_launches = 0;
}  // namespace fgs

extern "C" {

int fgs_abi_version(void) { return FGS_B200_ABI_VERSION; }
const char* fgs_last_error(void) { return fgs::g_err; }
int64_t fgs_launch_count(void) { return fgs::g_launches; }
void fgs_reset_launch_count(void) { fgs::g_launches = 0; }

int fgs_set_device(int device) {
  return fgs::check_cuda(cudaSetDevice(device), "cudaSetDevice");
}

int fgs_device_sm_count(int device, int* sm_count, int* cc_major, int* cc_minor) {
  cudaDeviceProp p;
  int rc = fgs::check_cuda(cudaGetDeviceProperties(&p, device), "cudaGetDeviceProperties");
  if (rc) return rc;
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  return 0;
}

}  // extern "C"
