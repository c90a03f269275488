// C-ABI of the fused fine-stage step (include/fgs_b200_fine.h): kernel sequencing only.
//   fgs_fine_forward  : march -> scan -> level1 -> scan -> compact2 -> feat -> 8 GEMMs -> composite
//   fgs_fine_loss     : composite with the loss block + first backward hop fused
//   fgs_fine_backward : 8 layers x (dW GEMM, db colsum, dX GEMM) -> feat_bwd -> level1_bwd
// plus the stand-alone DenseGrid / MaskCache trilinear ops.  No host synchronisation anywhere.
#include "fine_common.cuh"
#include "tc_api.cuh"

namespace fine {
int launch_march(const FgsFineParams& P, cudaStream_t st);
int launch_feat(const FgsFineParams& P, cudaStream_t st);
int launch_feat_bwd(const FgsFineParams& P, float weight_orientation, cudaStream_t st);
int launch_composite(const FgsFineParams& P, const FgsFineLoss* L, cudaStream_t st);
int launch_level1_bwd(const FgsFineParams& P, const float* g_last, cudaStream_t st);
}  // namespace fine

namespace {

using namespace fine;
constexpr int W = FGS_MLP_WIDTH;

int check_params(const FgsFineParams* p) {
  FGS_REQUIRE(p != nullptr, "params is NULL");
  FGS_REQUIRE(p->n_rays > 0 && p->n_rays <= (1 << 20), "n_rays out of range (1..2^20)");
  FGS_REQUIRE(p->X >= 2 && p->Y >= 2 && p->Z >= 2, "grid too small");
  FGS_REQUIRE(p->cap_words > 0 && p->cap_m1 > 0 && p->cap_m2 > 0, "capacities must be > 0");
  FGS_REQUIRE(p->stepdist > 0.f, "stepdist must be > 0");
  return 0;
}

float* hidden(const FgsFineParams& P, int i) { return P.H + (size_t)i * P.cap_m2 * W; }

uint32_t* hbits(const FgsFineParams& P, int i) { return P.hbits + (size_t)i * P.cap_m2 * 8; }

// forward step `layer`: C = act(A W_layer^T + b_layer) (straight planes)
ChainStepHost fwd_step(const FgsFineParams& P, const WSplitLayout& L, int layer, const float* A, int lda, int K,
                       float* C, int ldc, int N, int relu, int hid) {
  ChainStepHost s{};
  s.A = A; s.lda = lda; s.K = K;
  s.Bh = P.w_split + L.hi[layer]; s.Bl = P.w_split + L.lo[layer]; s.ldb = L.in_ld[layer];
  s.C = C; s.ldc = ldc; s.N = N; s.n0 = 0;
  s.bias = P.B[layer]; s.relu = relu;
  s.bits_in = nullptr; s.bits_out = (relu && hid >= 0) ? hbits(P, hid) : nullptr;
  return s;
}
// operand planes of the 8 weight matrices in one launch: straight TF32 hi / lo (forward), transposed bf16 hi / mid (backward)
int prep_weights(const FgsFineParams& P, cudaStream_t st) {
  FGS_REQUIRE(P.w_split != nullptr && P.hbits != nullptr, "w_split / hbits workspace is NULL");
  const WSplitLayout L = wsplit_layout();
  SplitJobHost jobs[2 * FGS_MLP_LAYERS];
  for (int l = 0; l < FGS_MLP_LAYERS; ++l) {
    jobs[2 * l] = {P.W[l], P.w_split + L.hi[l], P.w_split + L.lo[l], L.out[l], L.in_ld[l], L.in_ld[l], L.in_ld[l], 0, 0};
    jobs[2 * l + 1] = {P.W[l], P.w_split + L.thi[l], P.w_split + L.tlo[l], L.out[l], L.in_ld[l], L.in_ld[l], L.ldt[l], 1, 1};
  }
  return weight_split_jobs(jobs, 2 * FGS_MLP_LAYERS, st);
}

}  // namespace

extern "C" {

int64_t fgs_fine_wsplit_floats(void) { return (int64_t)wsplit_layout().total; }

int fgs_set_dw_sm_reserve(int n_sms) {
  FGS_REQUIRE(n_sms >= 0 && n_sms < FGS_NUM_SMS, "reserve must be in [0, number of SMs)");
  set_dw_sm_reserve(n_sms);
  return 0;
}

int fgs_fine_forward(const FgsFineParams* p, fgs_stream_t stream) {
  int rc = check_params(p);
  if (rc) return rc;
  const FgsFineParams& P = *p;
  cudaStream_t st = (cudaStream_t)stream;
  if ((rc = launch_march(P, st))) return rc;
  if ((rc = launch_feat(P, st))) return rc;
  const int* m2 = P.counters + FGS_CNT_M2;
  const int cap = P.cap_m2;
  if ((rc = prep_weights(P, st))) return rc;
  // rgbnet (model/nerf.py:125-134, 877) and refnet (:136-151, 883-884) as one persistent launch
  const WSplitLayout L = wsplit_layout();
  const ChainStepHost steps[8] = {
      fwd_step(P, L, 0, P.X0, FGS_X0_LD, FGS_X0_LD, hidden(P, 0), W, W, 1, 0),
      fwd_step(P, L, 1, hidden(P, 0), W, W, hidden(P, 1), W, W, 1, 1),
      fwd_step(P, L, 2, hidden(P, 1), W, W, hidden(P, 2), W, W, 1, 2),
      fwd_step(P, L, 3, hidden(P, 2), W, W, P.XR, FGS_XR_LD, W, 0, -1),
      fwd_step(P, L, 4, P.XR, FGS_XR_LD, FGS_XR_LD, hidden(P, 3), W, W, 1, 3),
      fwd_step(P, L, 5, hidden(P, 3), W, W, hidden(P, 4), W, W, 1, 4),
      fwd_step(P, L, 6, hidden(P, 4), W, W, hidden(P, 5), W, W, 1, 5),
      fwd_step(P, L, 7, hidden(P, 5), W, W, P.logit, 4, 3, 0, -1)};
  if ((rc = mlp_chain(steps, 8, cap, m2, st))) return rc;
  if (P.skip_composite) return 0;   // training: fgs_fine_loss runs the compositing with the loss fused
  return launch_composite(P, nullptr, st);
}

int fgs_fine_loss(const FgsFineParams* p, const FgsFineLoss* l, fgs_stream_t stream) {
  int rc = check_params(p);
  if (rc) return rc;
  FGS_REQUIRE(l && l->target && l->g_alphainv_last && l->loss_out, "loss descriptor incomplete");
  FGS_REQUIRE(l->n_rays_global >= p->n_rays, "n_rays_global < n_rays");
  // The MLP gradient bucket is accumulated into from here on (bias gradients: by the compositing kernel for the last
  // layer, by the backward chain's epilogues for the others; weights: by the stream-K weight-gradient kernel), so it
  // is cleared HERE unless the caller says the previous step's Adam already did (mlp_grads_zeroed).
  if (!p->mlp_grads_zeroed) {
    cudaStream_t st = (cudaStream_t)stream;
    if (p->g_mlp_flat && p->g_mlp_floats > 0) {
      rc = fgs::check_cuda(cudaMemsetAsync(p->g_mlp_flat, 0, sizeof(float) * (size_t)p->g_mlp_floats, st), "g_mlp memset");
      if (rc) return rc;
    } else {
      const int in_dim[8] = {FGS_X0_LD, W, W, W, FGS_XR_LD, W, W, W};
      const int out_dim[8] = {W, W, W, W, W, W, W, 3};
      for (int i = 0; i < 8; ++i) {
        if ((rc = fgs::check_cuda(cudaMemsetAsync(p->gW[i], 0, sizeof(float) * (size_t)in_dim[i] * out_dim[i], st), "gW memset"))) return rc;
        if ((rc = fgs::check_cuda(cudaMemsetAsync(p->gB[i], 0, sizeof(float) * (size_t)out_dim[i], st), "gB memset"))) return rc;
      }
    }
    fgs::prof_mark("memset", st);
  }
  return launch_composite(*p, l, (cudaStream_t)stream);
}

int fgs_fine_backward(const FgsFineParams* p, const FgsFineLoss* l, fgs_stream_t stream) {
  return fgs_fine_backward_phase(p, l, 0, stream);
}

int fgs_fine_backward_phase(const FgsFineParams* p, const FgsFineLoss* l, int phase, fgs_stream_t stream) {
  int rc = check_params(p);
  if (rc) return rc;
  FGS_REQUIRE(l && l->g_alphainv_last, "loss descriptor incomplete");
  FGS_REQUIRE(phase >= 0 && phase <= 2, "phase must be 0 (all), 1 (data path) or 2 (weight gradients)");
  FGS_REQUIRE(p->dZ != nullptr, "dZ workspace is NULL");
  const FgsFineParams& P = *p;
  cudaStream_t st = (cudaStream_t)stream;
  const int* m2 = P.counters + FGS_CNT_M2;
  const int cap = P.cap_m2;
  const int in_dim[8] = {FGS_X0_LD, W, W, W, FGS_XR_LD, W, W, W};
  const int out_dim[8] = {W, W, W, W, W, W, W, 3};
  {
    const WSplitLayout L = wsplit_layout();
    FGS_REQUIRE(P.dXRp && P.dlogit_p, "dXRp / dlogit_p workspaces are NULL");
    // gradient planes: layer slot i of dZ = [hi plane | mid plane], each [cap_m2][256] bf16 (dZ6, dZ5, dZ4, dZ2, dZ1, dZ0)
    const size_t plane = (size_t)P.cap_m2 * W;                    // elements per plane
    auto zhi = [&](int i) { return reinterpret_cast<uint16_t*>(P.dZ + (size_t)i * P.cap_m2 * W); };
    auto zmid = [&](int i) { return zhi(i) + plane; };
    uint16_t* xr_hi = P.dXRp; uint16_t* xr_mid = P.dXRp + plane;
    uint16_t* lg_hi = P.dlogit_p; uint16_t* lg_mid = P.dlogit_p + (size_t)P.cap_m2 * 8;
    if (phase != 2) {
      // ---- data path: backward-data chain in one launch on bf16 planes, dZ_l = (dZ_{l+1} W_{l+1}) masked by relu'(H_l);
      // the epilogue of the step that produces dZ_l also sums its columns (= the bias gradient of layer l).  Then the
      // scatters into the grids.  The weight gradients (phase 2) only read what this phase wrote, so a data-parallel
      // caller can exchange the grid gradients while they run. ----
      auto step = [&](int layer, const uint16_t* ahi, const uint16_t* amid, int lda, int K, uint16_t* chi, uint16_t* cmid,
                      float* c32, int ldc, int N, int n0, int hid, float* colsum) {
        BwStepHost s{};
        s.Ahi = ahi; s.Amid = amid; s.lda = lda; s.K = K;
        s.Bh = P.w_split + L.thi[layer]; s.Bl = P.w_split + L.tlo[layer]; s.ldb = L.ldt[layer];
        s.C = c32; s.ldc = ldc; s.Chi = chi; s.Cmid = cmid; s.out_planes = chi != nullptr;
        s.N = N; s.n0 = n0; s.bits_in = hid >= 0 ? hbits(P, hid) : nullptr; s.colsum = colsum;
        return s;
      };
      const BwStepHost steps[9] = {
          step(7, lg_hi, lg_mid, 8, 3, zhi(0), zmid(0), nullptr, 0, W, 0, 5, P.gB[6]),
          step(6, zhi(0), zmid(0), W, W, zhi(1), zmid(1), nullptr, 0, W, 0, 4, P.gB[5]),
          step(5, zhi(1), zmid(1), W, W, zhi(2), zmid(2), nullptr, 0, W, 0, 3, P.gB[4]),
          step(4, zhi(2), zmid(2), W, W, xr_hi, xr_mid, nullptr, 0, W, 0, -1, P.gB[3]),          // dXR[:, 0:256] = dZ3 (linear layer)
          step(4, zhi(2), zmid(2), W, W, nullptr, nullptr, P.dXR, FGS_XR_LD, FGS_XR_LD, 256, -1, nullptr),   // dXR[:, 256:320] -> feat_bwd
          step(3, xr_hi, xr_mid, W, W, zhi(3), zmid(3), nullptr, 0, W, 0, 2, P.gB[2]),
          step(2, zhi(3), zmid(3), W, W, zhi(4), zmid(4), nullptr, 0, W, 0, 1, P.gB[1]),
          step(1, zhi(4), zmid(4), W, W, zhi(5), zmid(5), nullptr, 0, W, 0, 0, P.gB[0]),
          step(0, zhi(5), zmid(5), W, W, nullptr, nullptr, P.dX0, FGS_X0_LD, FGS_X0_LD, 0, -1, nullptr)};
      if ((rc = mlp_bwd_planes(steps, 9, cap, m2, st))) return rc;
      if ((rc = launch_feat_bwd(P, l->weight_orientation, st))) return rc;
      if ((rc = launch_level1_bwd(P, l->g_alphainv_last, st))) return rc;
    }
    if (phase != 1) {
      // ---- weight gradients: dW_l = dZ_l^T X_l, all layers in one launch (3xBF16; dZ_l read as planes).  The bias
      // gradients were accumulated by fgs_fine_loss (layer 7) and by the chain above; the bucket was cleared before. ----
      const uint16_t* ahi[8] = {zhi(5), zhi(4), zhi(3), xr_hi, zhi(2), zhi(1), zhi(0), lg_hi};
      const uint16_t* amid[8] = {zmid(5), zmid(4), zmid(3), xr_mid, zmid(2), zmid(1), zmid(0), lg_mid};
      const int ldp[8] = {W, W, W, W, W, W, W, 8};
      const float* xin[8] = {P.X0, hidden(P, 0), hidden(P, 1), hidden(P, 2), P.XR, hidden(P, 3), hidden(P, 4), hidden(P, 5)};
      DwProblemHost dw[8];
      for (int i = 0; i < 8; ++i) {
        dw[i] = DwProblemHost{nullptr, xin[i], P.gW[i], nullptr, out_dim[i], in_dim[i], 0, in_dim[i], in_dim[i]};
        dw[i].Ahi = ahi[i]; dw[i].Amid = amid[i]; dw[i].ldp = ldp[i];
      }
      // the last layer has 3 out-features: a 256-row MMA tile would spend 12 % of the kernel on it; fp32 SIMT instead
      if ((rc = mlp_dw(dw, 7, cap, m2, st, true))) return rc;
      if ((rc = dw_small(out_dim[7], in_dim[7], cap, P.dlogit, 4, xin[7], in_dim[7], P.gW[7], in_dim[7], m2, st))) return rc;
    }
    return 0;
  }
}

}  // extern "C"

// ---- stand-alone trilinear ops (DenseGrid.forward/backward, MaskCache) ----------------------
namespace {

template <bool CL>
__global__ void __launch_bounds__(256)
grid_sample_fwd_kernel(const float* __restrict__ grid, int C, fine::GridGeo g, const float* __restrict__ xyz,
                       long long M, float* __restrict__ out) {
  for (long long p = blockIdx.x * 256LL + threadIdx.x; p < M; p += (long long)gridDim.x * 256) {
    float iw, ih, id;
    fine::cont_index(g, xyz[3 * p], xyz[3 * p + 1], xyz[3 * p + 2], iw, ih, id);
    const Tri t = tri_setup(iw, ih, id);
    const size_t V = (size_t)g.X * g.Y * g.Z;
    for (int ch = 0; ch < C; ++ch) {
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int x = t.x0 + TRI_DX(c), y = t.y0 + TRI_DY(c), z = t.z0 + TRI_DZ(c);
        if (x >= 0 && x < g.Z && y >= 0 && y < g.Y && z >= 0 && z < g.X) {
          const size_t vox = ((size_t)z * g.Y + y) * g.Z + x;
          acc = __fmaf_rn(__ldg(grid + (CL ? vox * C + ch : (size_t)ch * V + vox)), t.w[c], acc);
        }
      }
      out[p * C + ch] = acc;
    }
  }
}

template <bool CL>
__global__ void __launch_bounds__(256)
grid_sample_bwd_kernel(float* __restrict__ g_grid, int C, fine::GridGeo g, const float* __restrict__ xyz,
                       long long M, const float* __restrict__ g_out) {
  for (long long p = blockIdx.x * 256LL + threadIdx.x; p < M; p += (long long)gridDim.x * 256) {
    float iw, ih, id;
    fine::cont_index(g, xyz[3 * p], xyz[3 * p + 1], xyz[3 * p + 2], iw, ih, id);
    const Tri t = tri_setup(iw, ih, id);
    const size_t V = (size_t)g.X * g.Y * g.Z;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int x = t.x0 + TRI_DX(c), y = t.y0 + TRI_DY(c), z = t.z0 + TRI_DZ(c);
      if (x >= 0 && x < g.Z && y >= 0 && y < g.Y && z >= 0 && z < g.X) {
        const size_t vox = ((size_t)z * g.Y + y) * g.Z + x;
        for (int ch = 0; ch < C; ++ch)
          atomicAdd(g_grid + (CL ? vox * C + ch : (size_t)ch * V + vox), __fmul_rn(t.w[c], g_out[p * C + ch]));
      }
    }
  }
}

fine::GridGeo host_geo(const float* mn, const float* mx, int X, int Y, int Z) {
  fine::GridGeo g;
  for (int c = 0; c < 3; ++c) {
    g.mn[c] = mn[c];
    g.ext[c] = mx[c] - mn[c];  // fp32 subtraction, as torch evaluates (xyz_max - xyz_min)
  }
  g.X = X; g.Y = Y; g.Z = Z;
  return g;
}

}  // namespace

extern "C" {

// xyz_min / xyz_max are HOST pointers to 3 floats here (scalars of the grid object).
int fgs_grid_sample_fwd(const float* grid, int C, int X, int Y, int Z, int channels_last,
                        const float* xyz, int64_t M, const float* xyz_min, const float* xyz_max,
                        float* out, fgs_stream_t stream) {
  FGS_REQUIRE(C > 0 && X > 0 && Y > 0 && Z > 0 && M >= 0, "bad size");
  if (M == 0) return 0;
  const fine::GridGeo g = host_geo(xyz_min, xyz_max, X, Y, Z);
  const int grid_dim = fgs::stream_grid(M, 256, 8);
  if (channels_last)
    grid_sample_fwd_kernel<true><<<grid_dim, 256, 0, (cudaStream_t)stream>>>(grid, C, g, xyz, M, out);
  else
    grid_sample_fwd_kernel<false><<<grid_dim, 256, 0, (cudaStream_t)stream>>>(grid, C, g, xyz, M, out);
  return fgs::check_launch("grid_sample_fwd", (cudaStream_t)stream);
}

int fgs_grid_sample_bwd(float* g_grid, int C, int X, int Y, int Z, int channels_last,
                        const float* xyz, int64_t M, const float* xyz_min, const float* xyz_max,
                        const float* g_out, fgs_stream_t stream) {
  FGS_REQUIRE(C > 0 && X > 0 && Y > 0 && Z > 0 && M >= 0, "bad size");
  if (M == 0) return 0;
  const fine::GridGeo g = host_geo(xyz_min, xyz_max, X, Y, Z);
  const int grid_dim = fgs::stream_grid(M, 256, 8);
  if (channels_last)
    grid_sample_bwd_kernel<true><<<grid_dim, 256, 0, (cudaStream_t)stream>>>(g_grid, C, g, xyz, M, g_out);
  else
    grid_sample_bwd_kernel<false><<<grid_dim, 256, 0, (cudaStream_t)stream>>>(g_grid, C, g, xyz, M, g_out);
  return fgs::check_launch("grid_sample_bwd", (cudaStream_t)stream);
}

}  // extern "C"
