// fgs_fine_train_step: one single-GPU training iteration sequenced inside the library (include/fgs_b200_fine.h).
// Kernel sequencing only -- every launch goes through the public entry points the host-sequenced step uses, in the
// same order, so both drivers produce the same results.
#include "common.cuh"
#include "../../include/fgs_b200_fine.h"

extern "C" int fgs_fine_train_step(const FgsFineParams* p, const FgsFineLoss* l, const FgsFineStep* s,
                                   fgs_stream_t stream) {
  FGS_REQUIRE(p && l && s, "train_step: NULL descriptor");
  FGS_REQUIRE(l->loss_out, "train_step: loss_out is NULL");
  FGS_REQUIRE(s->sdf_m && s->sdf_v && s->k0_m && s->k0_v, "train_step: Adam state missing");
  FGS_REQUIRE(!s->sparse_k0 || (s->brick_list && p->k0_brick_flags), "train_step: sparse k0 needs brick flags + list");
  FGS_REQUIRE(!s->tv_smooth || s->tv_scratch, "train_step: tv_scratch missing");
  FGS_REQUIRE(s->n_mlp >= 0 && (s->n_mlp == 0 || (s->mlp_p && s->mlp_g && s->mlp_m && s->mlp_v && s->mlp_n &&
                                                   s->mlp_steps && s->mlp_lrs)),
              "train_step: MLP optimiser arrays missing");
  int rc = 0;
  const int X = p->X, Y = p->Y, Z = p->Z;
  const int64_t V = (int64_t)X * Y * Z;

  // ---- forward + loss + data-path backward (model/nerf_training.py:300-347) ----
  FgsFineParams P = *p;
  P.skip_composite = 1;                                  // fgs_fine_loss composites with the loss fused
  if ((rc = fgs_memset_zero(l->loss_out, 5 * sizeof(double), stream))) return rc;
  if ((rc = fgs_fine_forward(&P, stream))) return rc;
  if ((rc = fgs_fine_loss(&P, l, stream))) return rc;
  if ((rc = fgs_fine_backward_phase(&P, l, 1, stream))) return rc;
  // ---- regularisers on their schedule (:330-336, :353-360); full-grid stencils stay on the main stream ----
  if (s->tv_smooth &&
      (rc = fgs_smooth_grad_tv_add_grad(P.sdf, P.g_sdf, s->nonempty, s->tv_scratch, X, Y, Z, P.voxel_size,
                                        s->smooth_weight, s->smooth_n_sel, l->loss_out + 4, stream)))
    return rc;
  if (s->tv_sdf && (rc = fgs_total_variation_add_grad(P.sdf, P.g_sdf, nullptr, s->sdf_tv_weight, s->sdf_tv_weight,
                                                      s->sdf_tv_weight, 1, X, Y, Z, s->sdf_tv_dense, 0, stream)))
    return rc;
  // ---- the grid gradients are final: their optimiser kernels first (g_sdf is still L2-resident), then the weight
  // gradients and the Adam of the MLP tensors.  Everything on ONE stream: the optimiser kernels used to run on a side
  // stream "beside" the weight-gradient kernel, but its persistent CTAs hold every SM's shared memory, so the side
  // stream only started when they retired and the join then WAITED for it (measured: 0.13 ms exposed per step). ----
  if ((rc = fgs_adam_upd_zero_grad((float*)P.sdf, P.g_sdf, s->sdf_m, s->sdf_v, V, s->adam_step, s->beta1, s->beta2,
                                   s->lr_sdf, s->eps, stream)))
    return rc;
  if (s->sparse_k0) {
    if ((rc = fgs_adam_upd_brick_flags((float*)P.k0, P.g_k0, s->k0_m, s->k0_v, P.k0_brick_flags, s->brick_list, X, Y, Z, FGS_K0_DIM,
                                       s->adam_step, s->beta1, s->beta2, s->lr_k0, s->eps, stream)))
      return rc;
  } else {
    if ((rc = fgs_adam_upd((float*)P.k0, P.g_k0, s->k0_m, s->k0_v, nullptr, V * FGS_K0_DIM, s->adam_step, s->beta1,
                           s->beta2, s->lr_k0, s->eps, 1, stream)))
      return rc;
    if ((rc = fgs_memset_zero(P.g_k0, V * FGS_K0_DIM * (int64_t)sizeof(float), stream))) return rc;
    if (P.k0_brick_flags) {
      const int64_t nb = (int64_t)((X + 3) / 4) * ((Y + 3) / 4) * ((Z + 3) / 4);
      if ((rc = fgs_memset_zero(P.k0_brick_flags, nb, stream))) return rc;
    }
  }
  if ((rc = fgs_fine_backward_phase(&P, l, 2, stream))) return rc;
  if (s->n_mlp > 0 && (rc = fgs_adam_upd_multi_zero_grad(s->n_mlp, s->mlp_p, s->mlp_g, s->mlp_m, s->mlp_v, s->mlp_n, s->mlp_steps,
                                               s->mlp_lrs, s->beta1, s->beta2, s->eps, stream)))
    return rc;
  return 0;
}

// ---- the iteration as one CUDA graph ---------------------------------------------------------------------------------
namespace {
struct StepGraph {
  cudaGraphExec_t exec = nullptr;
};
StepGraph g_step_graphs[64][4];          // [device][tv_smooth | tv_sdf << 1]
}  // namespace

extern "C" int fgs_fine_train_step_graph(const FgsFineParams* p, const FgsFineLoss* l, const FgsFineStep* s,
                                         fgs_stream_t stream, int* used_graph) {
  if (used_graph) *used_graph = 0;
  FGS_REQUIRE(p && l && s, "train_step_graph: NULL descriptor");
  cudaStream_t st = (cudaStream_t)stream;
  int dev = 0;
  if (fgs::g_prof_on || st == nullptr || st == cudaStreamLegacy || cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64)
    return fgs_fine_train_step(p, l, s, stream);
  StepGraph& G = g_step_graphs[dev][(s->tv_smooth ? 1 : 0) | (s->tv_sdf ? 2 : 0)];
  int rc = 0;
  if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
    cudaGetLastError();                              // e.g. the caller is capturing this stream itself: join its capture
    return fgs_fine_train_step(p, l, s, stream);
  }
  const int64_t launches_before = fgs::g_launches;
  const int rc_step = fgs_fine_train_step(p, l, s, stream);
  cudaGraph_t graph = nullptr;
  cudaError_t e = cudaStreamEndCapture(st, &graph);
  if (rc_step) {                                     // the step's own error wins; the capture is void
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    return rc_step;
  }
  if (e != cudaSuccess || !graph) {
    cudaGetLastError();
    fgs::g_launches = launches_before;
    return fgs_fine_train_step(p, l, s, stream);     // capture not possible here: enqueue directly
  }
  if (G.exec) {
    cudaGraphExecUpdateResultInfo info;
    if (cudaGraphExecUpdate(G.exec, graph, &info) != cudaSuccess) {   // different topology (e.g. another row-range count)
      cudaGetLastError();
      cudaGraphExecDestroy(G.exec);
      G.exec = nullptr;
    }
  }
  if (!G.exec) {
    e = cudaGraphInstantiate(&G.exec, graph, 0);
    if (e != cudaSuccess) {
      cudaGraphDestroy(graph);
      G.exec = nullptr;
      return fgs::check_cuda(e, "graph instantiate");
    }
  }
  cudaGraphDestroy(graph);
  rc = fgs::check_cuda(cudaGraphLaunch(G.exec, st), "graph launch");
  if (!rc && used_graph) *used_graph = 1;
  return rc;
}
