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

// ---- per-launch timing (bench.py / profiles): one cudaEvent per launch on the launching stream
namespace fgs {
bool g_prof_on = false;          // per-launch timing active (fgs_prof_begin .. fgs_prof_end)
namespace {
constexpr int kMaxEv = 16384;
int g_prof_n = 0, g_prof_created = 0;
cudaEvent_t g_ev[kMaxEv];
const char* g_ev_name[kMaxEv];
cudaStream_t g_ev_stream[kMaxEv];
}  // namespace

void prof_mark(const char* what, cudaStream_t st) {
  if (!g_prof_on || g_prof_n >= kMaxEv) return;
  if (g_prof_n >= g_prof_created) {
    if (cudaEventCreate(&g_ev[g_prof_created]) != cudaSuccess) return;
    ++g_prof_created;
  }
  cudaEventRecord(g_ev[g_prof_n], st);
  g_ev_name[g_prof_n] = what;
  g_ev_stream[g_prof_n] = st;
  ++g_prof_n;
}
}  // namespace fgs

extern "C" {

int fgs_prof_begin(fgs_stream_t stream) {
  fgs::g_prof_n = 0;
  fgs::g_prof_on = true;
  fgs::prof_mark("begin", (cudaStream_t)stream);
  return 0;
}

int fgs_prof_end(void) {
  fgs::g_prof_on = false;
  if (fgs::g_prof_n == 0) return 0;
  int rc = fgs::check_cuda(cudaEventSynchronize(fgs::g_ev[fgs::g_prof_n - 1]), "prof sync");
  return rc ? -1 : fgs::g_prof_n - 1;
}

int fgs_prof_get(int i, const char** name, float* ms) {
  if (i < 0 || i + 1 >= fgs::g_prof_n) return FGS_E_BADARG;
  *name = fgs::g_ev_name[i + 1];
  // duration of launch i + 1 = time since the previous event on the SAME stream (callers that fork work onto
  // another stream mark the fork / join points, fgs_prof_mark(4 | 5), so every stream segment starts with one)
  int prev = i;
  for (int j = i; j >= 0; --j)
    if (fgs::g_ev_stream[j] == fgs::g_ev_stream[i + 1]) { prev = j; break; }
  return fgs::check_cuda(cudaEventElapsedTime(ms, fgs::g_ev[prev], fgs::g_ev[i + 1]), "prof elapsed");
}

// mark a non-library launch (e.g. an NCCL collective enqueued by the host) in the timing stream
int fgs_prof_mark(int what, fgs_stream_t stream) {
  static const char* names[] = {"other", "allreduce", "h2d copy", "d2h copy", "fork", "join"};
  fgs::prof_mark(names[(what >= 0 && what < 6) ? what : 0], (cudaStream_t)stream);
  return 0;
}

int fgs_memset_zero(void* ptr, int64_t bytes, fgs_stream_t stream) {
  if (bytes <= 0) return 0;
  int rc = fgs::check_cuda(cudaMemsetAsync(ptr, 0, (size_t)bytes, (cudaStream_t)stream), "memset");
  fgs::prof_mark("memset", (cudaStream_t)stream);
  return rc;
}

}  // extern "C"
