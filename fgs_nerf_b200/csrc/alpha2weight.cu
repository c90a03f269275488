// alpha -> weights: segmented transmittance scan keyed on the sorted ray_id (sm_100a).
//
// Replaces model/cuda/render_utils_kernel.cu:577-707 (alpha2weight_cuda_kernel,
// __set_i_for_segment_start_end, alpha2weight_backward_cuda_kernel).  The reference runs ONE
// THREAD per ray, serially over the ray's samples with uncoalesced loads, plus a host sync at
// kernel.cu:635.  Here a WARP owns a ray: lanes load 32 alphas with one coalesced request,
// the recurrence is then carried through registers with warp shuffles, and results are stored
// coalesced.  The recurrence keeps the reference's exact rounding sequence
//     T[i]=Tc; w[i]=Tc*a (fp32);  Tc = (float)((double)Tc * (1.0-(double)a));  stop if Tc<1e-3
// so the early-stop index -- and with it every id/mask downstream -- is bit-identical.
#include "common.cuh"

namespace {

constexpr int kBlock = 256;

// kernel.cu:607-617 plus the host-side `i_end[ray_id[n_pts-1]] = n_pts` of kernel.cu:635.
__global__ void seg_bounds_kernel(const long long* __restrict__ ray_id, long long n_pts,
                                  long long* __restrict__ i_start, long long* __restrict__ i_end) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_pts;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = ray_id[i];
    if (i > 0) {
      const long long rp = ray_id[i - 1];
      if (r != rp) {
        i_start[r] = i;
        i_end[rp] = i;
      }
    }
    if (i == n_pts - 1) i_end[r] = n_pts;
  }
}

__global__ void __launch_bounds__(kBlock)
alpha2weight_kernel(const float* __restrict__ alpha, int n_rays, float* __restrict__ weight,
                    float* __restrict__ T, float* __restrict__ alphainv_last,
                    const long long* __restrict__ i_start, long long* __restrict__ i_end) {
  const int wpb = kBlock / 32;
  const int lane = lane_id();
  for (int ray = blockIdx.x * wpb + (threadIdx.x >> 5); ray < n_rays; ray += gridDim.x * wpb) {
    const long long i_s = i_start[ray];
    const long long i_e = i_end[ray];
    float Tc = 1.f;
    long long stop = i_e;  // exclusive index after the last visited sample
    bool stopped = false;
    for (long long base = i_s; base < i_e; base += 32) {
      const long long idx = base + lane;
      const float a = (idx < i_e) ? alpha[idx] : 0.f;
      float myT = 1.f, myW = 0.f;
      if (!stopped) {
        const int cnt = (int)min((long long)32, i_e - base);
        for (int j = 0; j < cnt; ++j) {
          const float aj = __shfl_sync(0xffffffffu, a, j);
          if (j == lane) {
            myT = Tc;
            myW = __fmul_rn(Tc, aj);
          }
          Tc = (float)__dmul_rn(__dsub_rn(1., (double)aj), (double)Tc);
          if ((double)Tc < 1e-3) {
            stopped = true;
            stop = base + j + 1;
            if (lane > j) { myT = 1.f; myW = 0.f; }
            break;
          }
        }
      }
      if (idx < i_e) {
        T[idx] = myT;
        weight[idx] = myW;
      }
    }
    if (lane == 0) {
      i_end[ray] = stop;
      alphainv_last[ray] = Tc;
    }
  }
}

// kernel.cu:654-677.  back_cum is an fp32 FMA chain (exactly the reference's order); each
// lane then forms its own gradient with the reference's double-precision divide.
__global__ void __launch_bounds__(kBlock)
alpha2weight_backward_kernel(const float* __restrict__ alpha, const float* __restrict__ weight,
                             const float* __restrict__ T, const float* __restrict__ alphainv_last,
                             const long long* __restrict__ i_start,
                             const long long* __restrict__ i_end, int n_rays,
                             const float* __restrict__ grad_weights,
                             const float* __restrict__ grad_last, float* __restrict__ grad) {
  const int wpb = kBlock / 32;
  const int lane = lane_id();
  for (int ray = blockIdx.x * wpb + (threadIdx.x >> 5); ray < n_rays; ray += gridDim.x * wpb) {
    const long long i_s = i_start[ray];
    const long long i_e = i_end[ray];
    float back = __fmul_rn(grad_last[ray], alphainv_last[ray]);
    // walk chunks from the far end; lane l of a chunk owns index top-1-l
    for (long long top = i_e; top > i_s; top -= 32) {
      const long long idx = top - 1 - lane;
      const bool valid = idx >= i_s;
      const float gw = valid ? grad_weights[idx] : 0.f;
      const float w = valid ? weight[idx] : 0.f;
      const float t = valid ? T[idx] : 0.f;
      const float a = valid ? alpha[idx] : 0.f;
      const int cnt = (int)min((long long)32, top - i_s);
      float myBack = 0.f;
      for (int j = 0; j < cnt; ++j) {
        if (j == lane) myBack = back;
        const float gwj = __shfl_sync(0xffffffffu, gw, j);
        const float wj = __shfl_sync(0xffffffffu, w, j);
        back = __fmaf_rn(gwj, wj, back);
      }
      if (valid) {
        const double den = __dadd_rn((double)__fsub_rn(1.f, a), 1e-10);
        grad[idx] = (float)__dsub_rn((double)__fmul_rn(gw, t), __ddiv_rn((double)myBack, den));
      }
    }
  }
}

}  // namespace

extern "C" {

int fgs_alpha2weight(const float* alpha, const int64_t* ray_id, int64_t n_pts, int n_rays,
                     float* weight, float* T, float* alphainv_last, int64_t* i_start,
                     int64_t* i_end, fgs_stream_t stream) {
  FGS_REQUIRE(n_pts >= 0 && n_rays >= 0, "negative size");
  cudaStream_t st = (cudaStream_t)stream;
  if (n_rays == 0) return 0;
  int rc = fgs::check_cuda(cudaMemsetAsync(i_start, 0, 8 * (size_t)n_rays, st), "a2w memset");
  if (rc) return rc;
  rc = fgs::check_cuda(cudaMemsetAsync(i_end, 0, 8 * (size_t)n_rays, st), "a2w memset");
  if (rc) return rc;
  if (n_pts > 0) {
    seg_bounds_kernel<<<fgs::stream_grid(n_pts, kBlock), kBlock, 0, st>>>(
        (const long long*)ray_id, n_pts, (long long*)i_start, (long long*)i_end);
    rc = fgs::check_launch("seg_bounds");
    if (rc) return rc;
  }
  alpha2weight_kernel<<<fgs::stream_grid(n_rays, kBlock / 32, 16), kBlock, 0, st>>>(
      alpha, n_rays, weight, T, alphainv_last, (const long long*)i_start, (long long*)i_end);
  return fgs::check_launch("alpha2weight");
}

int fgs_alpha2weight_backward(const float* alpha, const float* weight, const float* T,
                              const float* alphainv_last, const int64_t* i_start,
                              const int64_t* i_end, int n_rays, int64_t n_pts,
                              const float* grad_weights, const float* grad_last, float* grad,
                              fgs_stream_t stream) {
  FGS_REQUIRE(n_pts >= 0 && n_rays >= 0, "negative size");
  cudaStream_t st = (cudaStream_t)stream;
  if (n_pts == 0) return 0;
  int rc = fgs::check_cuda(cudaMemsetAsync(grad, 0, 4 * (size_t)n_pts, st), "a2w_bwd memset");
  if (rc || n_rays == 0) return rc;
  alpha2weight_backward_kernel<<<fgs::stream_grid(n_rays, kBlock / 32, 16), kBlock, 0, st>>>(
      alpha, weight, T, alphainv_last, (const long long*)i_start, (const long long*)i_end, n_rays,
      grad_weights, grad_last, grad);
  return fgs::check_launch("alpha2weight_backward");
}

}  // extern "C"
