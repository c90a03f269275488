/*
 * fgs_b200_fine.h -- C-ABI of the FUSED fine-stage step (new entry points; the reference has no
 * single call for this, it is the body of its training iteration):
 *   forward  = nerf.forward_fine                         (model/nerf.py:776-941)
 *   loss     = the loss block of the training loop       (model/nerf_training.py:307-345)
 *   backward = what autograd does for that graph         
 *   reg/opt  = sdf TV add-grad + smooth-grad TV + MaskedAdam (nerf_training.py:330-373)
 *
 * All memory is owned by the caller (the Python host allocates it as torch tensors); the
 * struct below only carries pointers and sizes.  Every launcher enqueues a fixed sequence of
 * kernels on `stream` with NO host synchronisation: data-dependent sample counts stay on the
 * device (`counters`), buffers are sized by capacity, so a whole step is CUDA-graph capturable.
 *
 * Sample levels (flat, ray-major, near->far inside a ray -- the reference's packing):
 *   level 1 = samples inside the bbox that pass the MaskCache test      (nerf.py:780-789)
 *   level a = level-1 samples with alpha > fast_color_thres            (nerf.py:802-810)
 *   level 2 = level-a samples visited before the T<1e-3 stop with weight > fast_color_thres
 *             (nerf.py:823-833) -- the only ones that reach the MLPs.
 * Levels 1 and a share the level-1 arrays (flag1 tells them apart); level 2 is compacted.
 *
 * Layouts:
 *   sdf   [X][Y][Z] fp32 (the reference's [1,1,X,Y,Z]);
 *   k0    [X][Y][Z][C] fp32 = torch.channels_last_3d storage of the reference's [1,C,X,Y,Z]
 *         (one voxel's 12 channels are three 128-bit loads);
 *   mask  [mX][mY][mZ] fp32 (MaskCache.sdf_mask after its max-pool, model/nerf.py:1199);
 *   MLP weights: nn.Linear layout W[out][in], bias[out]; layers 0-3 = rgbnet, 4-7 = refnet; the two
 *         input layers are stored zero-padded to the activation row length: W[0] is [256][112]
 *         (106 used), W[4] is [256][320] (307 used), so every GEMM operand row is 16-byte aligned
 *         (pad columns stay exactly 0 under Adam: their gradient is 0);
 *   activations: SAMPLE-MAJOR rows [cap_m2][ld] (K contiguous: the operand layout of the GEMMs).
 */
#ifndef FGS_B200_FINE_H
#define FGS_B200_FINE_H

#include <stdint.h>

#include "fgs_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define FGS_MLP_LAYERS 8
#define FGS_MLP_WIDTH 256
#define FGS_X0_DIM 106   /* rgbnet input (model/nerf.py:108-112, 874)                 */
#define FGS_X0_LD 112    /* padded row length of X0 (pad columns are written as 0)    */
#define FGS_XR_DIM 307   /* refnet input = 256 rgb_feat + 51 reflect_emb (nerf.py:883) */
#define FGS_XR_LD 320
#define FGS_K0_DIM 12

/* X0 column map (concat order of model/nerf.py:874) */
#define FGS_X0_K0 0      /* 12: k0                                   */
#define FGS_X0_XYZ 12    /* 33: xyz, sin(xyz*2^k), cos(xyz*2^k) k<5  */
#define FGS_X0_VIEW 45   /* 21: viewdirs emb, k<3                    */
#define FGS_X0_SDF 66    /*  1: sdf                                  */
#define FGS_X0_FEAT 67   /* 24: all_feat [tap(z-,z+,y-,y+,x-,x+)][d] */
#define FGS_X0_AGRAD 91  /* 12: all_grad [axis(z,y,x)][d] normalised */
#define FGS_X0_GRAD 103  /*  3: gradient (x,y,z)                     */

/* counters[] (device int32) */
#define FGS_CNT_M1 0        /* level-1 samples                                          */
#define FGS_CNT_M2 1        /* level-2 samples                                          */
#define FGS_CNT_OVERFLOW 2  /* bit0: steps/ray > bitmap, bit1: M1 > cap_m1, bit2: M2 > cap_m2 */
#define FGS_CNT_MA 3        /* level-a samples (alpha > thres, visited or not)          */
#define FGS_CNT_TOTAL_LO 4  /* total_len = sum of N_steps, low / high 32 bits           */
#define FGS_CNT_TOTAL_HI 5
#define FGS_CNT_M0 6        /* samples inside the bbox (before the MaskCache)           */
#define FGS_CNT_STICKY 7    /* OR of every overflow bit since the HOST last cleared it: fgs_fine_forward clears
                               counters[0..6] only, so an overflow in any iteration stays visible until read */
#define FGS_NUM_COUNTERS 8

/* flag1[] bits */
#define FGS_F_ALPHA 1   /* alpha > thres (level a)                   */
#define FGS_F_VISIT 2   /* reached before the early stop             */
#define FGS_F_L2 4      /* weight > thres (level 2)                  */

typedef struct FgsFineParams {
  /* ---- sizes ---- */
  int32_t n_rays;
  int32_t X, Y, Z;
  int32_t mX, mY, mZ;
  int32_t cap_words;            /* 32-bit words of the per-ray step bitmap */
  int32_t cap_m1, cap_m2;       /* sample capacities; cap_m2 % 128 == 0   */
  int32_t render_depth, render_normal;
  int32_t skip_composite;       /* fgs_fine_forward stops after the MLPs (fgs_fine_loss composites) */
  int32_t mlp_grads_zeroed;     /* the MLP gradient bucket is already all-zero (cleared by fgs_adam_upd_multi_zero_grad):
                                   fgs_fine_backward skips its memset */
  /* ---- scalars ---- */
  float xyz_min[3], xyz_max[3], mask_xyz_min[3], mask_xyz_max[3];
  float near_, far_, stepdist, dist, voxel_size, mask_thres, fast_color_thres, inv_s, bg;
  float pad0_;
  /* ---- inputs ---- */
  const float *rays_o, *rays_d, *viewdirs;       /* [N,3] */
  /* ---- parameters ---- */
  const float *sdf, *k0, *mask_grid;
  const float *W[FGS_MLP_LAYERS], *B[FGS_MLP_LAYERS];
  /* ---- per-ray workspace ---- */
  float *t_min, *t_max;                          /* [N] */
  int32_t *n_steps, *cnt1, *cnt2, *off1, *off2;  /* [N]; off*: [N+1] */
  uint32_t *bits;                                /* [N][cap_words] level-1 membership per step */
  float *alphainv_last;                          /* [N] */
  /* ---- level-1 samples, SoA, capacity cap_m1 ---- */
  int32_t *ray_id1, *step_id1;
  uint8_t *flag1;
  float *sdf1, *grad1 /*[3][cap_m1] x,y,z*/, *alpha1, *w1, *T1;
  float *d_sdf1, *d_grad1 /*[3][cap_m1]*/, *d_w1;  /* backward, valid where flag1 & FGS_F_L2 */
  /* ---- level-2 samples, capacity cap_m2 ---- */
  int32_t *idx1_of2, *ray_id2;
  float *X0;       /* [cap_m2][FGS_X0_LD]   rgbnet input                                   */
  float *H;        /* [6][cap_m2][256] hidden activations (post-ReLU): rgbnet 0-2, refnet 3-5 */
  float *XR;       /* [cap_m2][FGS_XR_LD]  cols 0..255 rgbnet output, 256..306 reflect emb  */
  float *logit;    /* [cap_m2][4] refnet output (pre-sigmoid), col 3 unused                 */
  float *normal2;  /* [cap_m2][4] normal (nerf.py:835), col 3 = n.v                         */
  /* ---- per-ray outputs ---- */
  float *rgb_marched, *sigmoid_rgb, *cum_weights, *depth, *normal_marched; /* [N,3],[N,3],[N],[N],[N,3] */
  int32_t *counters;                             /* [FGS_NUM_COUNTERS] device */
  /* ---- backward: parameter gradients (accumulated into; caller zeroes g_sdf, g_k0) ---- */
  float *g_sdf, *g_k0;
  float *gW[FGS_MLP_LAYERS], *gB[FGS_MLP_LAYERS]; /* overwritten */
  /* ---- backward scratch, capacity cap_m2 ---- */
  float *dA, *dB;  /* [cap_m2][256] ping-pong  */
  float *dXR;      /* [cap_m2][FGS_XR_LD]      */
  float *dX0;      /* [cap_m2][FGS_X0_LD]      */
  float *dlogit;   /* [cap_m2][4]              */
  uint16_t *dlogit_p;  /* [2][cap_m2][8] bf16 hi / mid planes of dlogit (written with it by fgs_fine_loss) */
  uint16_t *dXRp;      /* [2][cap_m2][256] bf16 hi / mid planes of dXR[:, 0:256] = the gradient at rgbnet's output */
  /* ---- tensor-core MLP workspace ---- */
  float *w_split;    /* fgs_fine_wsplit_floats() floats: hi / lo TF32 planes of the 8 weight matrices,
                        straight and transposed; rewritten by every fgs_fine_forward; zero-initialised */
  uint32_t *hbits;   /* [6][cap_m2][8] ReLU masks of the hidden activations, 1 bit per unit */
  float *g_mlp_flat;   /* optional: one buffer that contains every gW[i] / gB[i] (cleared with one memset) */
  int64_t g_mlp_floats;
  uint8_t *k0_brick_flags; /* [ceil(X/4)*ceil(Y/4)*ceil(Z/4)] set to 1 by the backward scatter for every
                        4x4x4 brick of k0 that received a gradient (may be NULL); see fgs_adam_upd_bricks */
  float *dZ;         /* [6][cap_m2][256] x 4 bytes: the pre-activation gradient of every hidden layer, each as TWO bf16
                        planes (hi [cap_m2][256], then mid [cap_m2][256]: the 3xBF16 operand format), written once by
                        the backward chain's epilogue and read by TMA by its next step and by the weight-gradient
                        kernel.  Zero-initialise it once: rows beyond the sample count are multiplied by zeros */
  uint8_t *k0_vox_flags; /* optional [X*Y*Z]: set to 1 by the backward scatter for every k0 voxel that received a
                        gradient (data-parallel exchange at voxel granularity, fgs_voxel_list); NULL = not recorded */
  const uint32_t *mask_occ; /* optional: fgs_mask_occupancy() of mask_grid -- one bit per CELL of the MaskCache grid, set
                        iff any of the cell's 8 corners is non-zero.  A sample in a cell whose corners are all 0
                        interpolates to exactly 0 < mask_thres (requires mask_thres > 0), so the march kernel skips its
                        8 gathers: the result is the same bit for bit.  NULL = always gather */
  uint8_t *sdf_brick_flags; /* optional [ceil(X/4)*ceil(Y/4)*ceil(Z/4)]: set to 1 by the backward scatters for every 4^3
                        brick of the SDF grid that received a gradient (data-parallel exchange: g_sdf then travels as
                        packed bricks like g_k0 instead of as the dense grid); NULL = not recorded */
  /* geometry_searching stage (model/nerf.py:962-967): mask_grid may be NULL (no MaskCache: every in-bbox sample passes),
   * and inc_mask, when given, is the MaskGrid of Incremental Voxel Training -- a sample survives iff
   * inc_mask[round(x * scale + shift)] != 0 (grid.py:274-284 = render_utils_cuda.maskcache_lookup) */
  const uint8_t *inc_mask;      /* optional [iX][iY][iZ] bytes */
  float inc_scale[3], inc_shift[3];
  int32_t iX, iY, iZ;
  int32_t pad2_;
} FgsFineParams;

/* words of FgsFineParams::mask_occ for an [mX][mY][mZ] mask grid, and the kernel that fills it (re-run whenever the
 * mask grid changes; model/nerf.py:1192-1209 builds the grid once per stage) */
int64_t fgs_mask_occupancy_words(int mX, int mY, int mZ);
int fgs_mask_occupancy(const float* mask_grid, int mX, int mY, int mZ, uint32_t* occ, fgs_stream_t stream);

/* size of FgsFineParams::w_split in floats */
int64_t fgs_fine_wsplit_floats(void);

/* forward_fine: sampling -> MaskCache -> SDF/gradient/alpha -> weights -> features -> rgbnet ->
 * refnet -> compositing.  Fills the per-ray outputs, counters and every saved intermediate. */
int fgs_fine_forward(const FgsFineParams* p, fgs_stream_t stream);

/* Loss block of model/nerf_training.py:307-327 (+ nerf.orientation_loss, nerf.py:469-478)
 * fused with the start of the backward pass: per ray it forms d loss / d rgb_marched,
 * d sigmoid_rgb, d alphainv_last, and pushes them through the compositing sums
 * (nerf.py:888-903) into dlogit [cap_m2][4] and d_w1.  The four scalar loss terms are ADDED
 * into loss_out[0..3] (mse, entropy_last, orientation, sigmoid_rgb; caller zeroes). */
typedef struct FgsFineLoss {
  const float* target;          /* [N,3] */
  float weight_main, weight_entropy_last, weight_orientation, sigmoid_rgb_loss;
  int32_t n_rays_global;        /* N of the whole (multi-GPU) batch: mse means divide by 3*this */
  int32_t has_last_ray;         /* this rank owns the global last ray (entropy quirk)          */
  float* g_alphainv_last;       /* [N] written                                                 */
  double* loss_out;             /* [5] device: mse (+ rgbper), entropy, orientation, sigmoid_rgb, smooth TV */
  float weight_rgbper;          /* coarse stage (model/nerf_training.py:311-314): per-point colour loss; 0 in the fine stage */
  int32_t rgbper_slot;          /* its value is added to loss_out[rgbper_slot] (0: into the mse term; the coarse engine
                                   keeps it apart in slot 6 so that PSNR can be formed from the mse alone) */
} FgsFineLoss;
int fgs_fine_loss(const FgsFineParams* p, const FgsFineLoss* l, fgs_stream_t stream);

/* Backward of fgs_fine_forward given what fgs_fine_loss left behind; accumulates into g_sdf,
 * g_k0 (atomics) and overwrites gW/gB. */
int fgs_fine_backward(const FgsFineParams* p, const FgsFineLoss* l, fgs_stream_t stream);

/* The same in two phases for data-parallel callers: phase 1 = everything that produces g_sdf / g_k0 (the
 * backward-data chain and the grid scatters), phase 2 = the weight / bias gradients gW / gB, which only read
 * what phase 1 left behind -- so the grid-gradient exchange can run while phase 2 does.  phase 0 = both. */
int fgs_fine_backward_phase(const FgsFineParams* p, const FgsFineLoss* l, int phase, fgs_stream_t stream);

/* Leave n_sms SMs free while the weight-gradient kernel of phase 2 runs (room for the collectives of an
 * overlapping gradient exchange); 0 = use the whole GPU (default). */
int fgs_set_dw_sm_reserve(int n_sms);

/* smooth_grad_tv term of nerf.density_total_variation (model/nerf.py:438-446) differentiated
 * analytically: g_sdf += d/dsdf [ weight * mean_{nonempty,3ch}( (conv3(G).detach() - G)^2 ) ],
 * G = central-difference gradient grid of the SDF (nerf.py:491-494).  `scratch` holds
 * 6*X*Y*Z floats; nonempty may be NULL (all voxels); n_sel = number of selected elements
 * (3 * masked voxels).  loss_out (may be NULL, double) has the loss value ADDED. */
int fgs_smooth_grad_tv_add_grad(const float* sdf, float* g_sdf, const uint8_t* nonempty,
                                float* scratch, int X, int Y, int Z, float voxel_size,
                                float weight, double n_sel, double* loss_out,
                                fgs_stream_t stream);

/* ------------------------------------------------------------------------------------
 * One whole single-GPU training iteration (model/nerf_training.py:300-373) sequenced by the library: the host makes
 * ONE call per iteration instead of ~20.  Measured on a B200 (profiles/r01_stage_ops_v2.json): the host needs 0.10 ms
 * to enqueue an iteration instead of 0.18 ms from Python; the device time of the iteration is unchanged (the step is
 * GPU-bound either way) -- this buys host headroom, not throughput.
 * Same launches and order as the host-sequenced step, all on `stream`:
 *   loss buffer clear, fgs_fine_forward (composite skipped), fgs_fine_loss, backward phase 1 (data path + grid
 *   scatters), smooth-gradient TV and SDF TV add-grad on their schedule, dense SDF Adam + gradient clear, unordered
 *   touched-brick list + brick Adam of k0, backward phase 2 (weight gradients), multi-tensor Adam of the 16 MLP tensors
 *   (which clears the gradients it consumes).
 * (Round 1 ran the grid optimiser kernels on `side_stream` beside the weight-gradient kernel; its persistent CTAs
 * hold every SM's shared memory, so the side stream only started when they retired and the join waited for it:
 * 0.13 ms exposed per step.  `side_stream` is kept in the descriptor for ABI stability and is unused.)
 * Everything is enqueued; nothing synchronises.
 * ------------------------------------------------------------------------------------ */
typedef struct FgsFineStep {
  /* regularisers of this iteration (host decides from its schedule) */
  int tv_smooth;            /* != 0: fgs_smooth_grad_tv_add_grad */
  float smooth_weight;      /* weight_tv_density * smooth_grad_tv */
  double smooth_n_sel;      /* 3 * selected voxels */
  const uint8_t* nonempty;  /* [X*Y*Z] or NULL */
  float* tv_scratch;        /* 6 * X*Y*Z floats (only read when tv_smooth) */
  int tv_sdf;               /* != 0: fgs_total_variation_add_grad on the SDF grid */
  float sdf_tv_weight;      /* already scaled by max(world_size)/128 and 1/n_rays_global */
  int sdf_tv_dense;
  /* optimiser */
  int adam_step;
  float beta1, beta2, eps;
  float lr_sdf, lr_k0;
  float *sdf_m, *sdf_v, *k0_m, *k0_v;
  int sparse_k0;            /* != 0: brick list + brick Adam; 0: dense masked Adam + clears */
  int32_t* brick_list;
  int n_mlp;                /* tensors of the multi-tensor Adam (HOST arrays below, n_mlp entries each) */
  float* const* mlp_p;
  float* const* mlp_g;            /* cleared by the step's Adam once consumed */
  float* const* mlp_m;
  float* const* mlp_v;
  const int64_t* mlp_n;
  const int* mlp_steps;
  const float* mlp_lrs;
  fgs_stream_t side_stream;
} FgsFineStep;

int fgs_fine_train_step(const FgsFineParams* p, const FgsFineLoss* l, const FgsFineStep* s, fgs_stream_t stream);

/* The same iteration launched as ONE CUDA graph: the call stream-captures fgs_fine_train_step (thread-local capture
 * mode), updates the cached executable graph of the same shape in place (cudaGraphExecUpdate: the per-iteration scalars
 * -- inv_s, Adam step, learning rates, ray pointers -- are kernel parameters of the captured nodes), and launches it.
 * One executable graph per step shape (TV regularisers on / off) and device.  `stream` must not be the legacy default
 * stream (capture is illegal there).  While per-launch timing is active (fgs_prof_begin) the iteration is enqueued
 * directly instead.  *used_graph (may be NULL) reports which way it went. */
int fgs_fine_train_step_graph(const FgsFineParams* p, const FgsFineLoss* l, const FgsFineStep* s, fgs_stream_t stream,
                              int* used_graph);

/* DenseGrid.forward / backward (model/grid.py:49-59) and MaskCache.forward (model/nerf.py:
 * 1202-1209) as stand-alone ops: ATen grid_sample arithmetic, any channel count.
 * channels_last: 0 = [C][X][Y][Z], 1 = [X][Y][Z][C].  out: [M][C].
 * xyz_min / xyz_max are HOST pointers to 3 floats (scalars of the grid object). */
int fgs_grid_sample_fwd(const float* grid, int C, int X, int Y, int Z, int channels_last,
                        const float* xyz, int64_t M, const float* xyz_min, const float* xyz_max,
                        float* out, fgs_stream_t stream);
int fgs_grid_sample_bwd(float* g_grid, int C, int X, int Y, int Z, int channels_last,
                        const float* xyz, int64_t M, const float* xyz_min, const float* xyz_max,
                        const float* g_out, fgs_stream_t stream);

/* Row-major fp32 GEMMs of the MLPs on the tcgen05 tensor cores (3xTF32 split, fp32-grade accuracy), exposed
 * stand-alone for tests and for callers with other layer shapes:
 *   mode 0 (NT): C[M,N] = A[M,K] * B[N,K]^T     mode 1 (NN): C = A[M,K] * B[K,N]
 * bias (may be NULL) is added per output column; relu != 0 applies max(0, .).
 * The row count of the sample dimension is read on the device from *rows_dev when rows_dev != NULL (clamped to
 * the host value).  Operands must be 16-byte aligned with leading dimensions % 4 == 0. */
/* The weight operand B is split once into TF32 hi / lo planes (internal scratch) and streamed by TMA; whole-K
 * accumulation in TMEM (error ~3e-6 of max|out|).  bits_out (mode 0, relu) receives the ReLU mask as [M][8]
 * words; bits_in (may be NULL) masks the result. */
int fgs_gemm_tcw(int mode, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
                 float* C, int ldc, const float* bias, int relu, const uint32_t* bits_in,
                 uint32_t* bits_out, const int32_t* rows_dev, fgs_stream_t stream);

/* C[M,N] += A[K,M]^T B[K,N] (3xTF32, split over K, atomically accumulated) with colsum[m] += sum_k A[k][m]
 * (may be NULL): the kernel the fused step runs for the weight / bias gradients. */
int fgs_gemm_tn(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                float* colsum, const int32_t* rows_dev, fgs_stream_t stream);

/* The kernels of the fused step's BACKWARD, stand-alone: operands as bf16 hi / mid planes, products
 * hi*hi + hi*mid + mid*hi on tcgen05 kind::f16 (3xBF16: ~1e-5 relative, half the tensor-core time of 3xTF32; the
 * forward keeps 3xTF32 so that the ReLU branch assignment stays the fp32 one).  Inside the fused step the planes are
 * written by the kernels that compute the values; these entries split their fp32 arguments in a pre-pass.
 * fgs_gemm_tcw_bf16: one step of the planes chain, C[M,N<=256] = A B^T (mode 0) / A B (mode 1), masked by bits_in
 *   when given, colsum[n] += column sums of the (masked) result when given; via_planes != 0 routes the result through
 *   the bf16-plane output path of the epilogue (C = hi + mid: the value the next step would read).
 * fgs_gemm_tn_bf16: the weight-gradient product C[M<=256,N] += A[K,M]^T B[K,N], colsum as fgs_gemm_tn. */
int fgs_gemm_tcw_bf16(int mode, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
                      float* C, int ldc, const uint32_t* bits_in, float* colsum, int via_planes,
                      const int32_t* rows_dev, fgs_stream_t stream);
int fgs_gemm_tn_bf16(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                     float* colsum, const int32_t* rows_dev, fgs_stream_t stream);

/* Debug aid: 20 x int64 device buffer (slots 10..19: the same phases of the weight-gradient kernel) that subsequent pre-split GEMM launches fill with the clock cycles
 * the roles of CTA (0,0) spent per phase (consumer drain / epilogue, producer wait / work, issuer waits on
 * a_full / b_full / r_empty / free, issuer total; tools/gemm_phase_probe.py).  NULL = off. */
int fgs_debug_gemm_phases(void* buf);

/* Timing aid: run only row part `task` (0 k0, 1 encodings, 2 reflection chain, 3..6 displacement levels) of the
 * level-2 feature kernels; -1 = all (default).  Results are incomplete while it is set. */
int fgs_debug_feat_task(int task);

/* ---- coarse stage (nerf.forward_coarse, model/nerf.py:943-1075) -------------------------------
 * Same sampling, MaskCache, NeuS alpha, compositing, loss and tensor-core MLP kernels as the fine stage; what differs:
 *   * the SDF value is sampled from the Gaussian-smoothed grid of THIS forward (fgs_conv3d_replicate, nerf.py:969),
 *     the gradient trilinearly from the central-difference grid of the RAW sdf (fgs_sdf_gradient_grid, :972-973);
 *   * alpha2weight runs twice: over all level-1 samples, ONE prune `weights > fast_color_thres`, and again over the
 *     pruned samples (:978-990) -- one in-order warp sweep per ray carries both recurrences;
 *   * features [k0 | xyz_emb | reflect_emb | normal | viewdirs_emb] -> refnet in -> width -> width -> 3 (:992-1006);
 *   * the loss adds weight_rgbper (FgsFineLoss; fgs_fine_loss is the loss entry of both stages); the SDF gradient
 *     flows back through the two grids: fgs_conv3d_replicate_bwd and fgs_sdf_gradient_grid_bwd add into g_sdf.
 * FgsFineParams supplies the shared workspaces, with these conventions: X0 holds the refnet input with row length
 * FgsCoarse.in_ld, H the two hidden layers, dZ / hbits two slots; the refnet's three layers live in slots 0, 1 and 7
 * of W / B / gW / gB, stored ZERO-PADDED to the kernels' 256-wide tiles: W[0] [256][in_ld], W[1] [256][256],
 * W[7] [3][256], B[0], B[1] [256] (rows / columns beyond refnet_width stay exactly 0 under Adam: their pre-activations
 * are 0, relu'(0) = 0, so every gradient that touches them is 0); w_split has fgs_coarse_wsplit_floats(in_ld) floats.
 * flag1: FGS_F_ALPHA on every level-1 sample (no alpha prune in this stage), FGS_F_L2 = pruned set (reaches the
 * refnet), FGS_F_VISIT = visited by the second recurrence before its T < 1e-3 stop; w1 / T1 are the SECOND pass's. */
typedef struct FgsCoarse {
  const float* sdf_smooth;   /* [X][Y][Z]    smooth_conv(sdf) of this forward                                   */
  const float* grad_grid;    /* [3][X][Y][Z] neus_sdf_gradient('interpolate'): d/dx, d/dy, d/dz                  */
  float* g_smooth;           /* [X][Y][Z]    gradient wrt sdf_smooth, accumulated into (caller zeroes)          */
  float* g_grad_grid;        /* [3][X][Y][Z] gradient wrt grad_grid, accumulated into (caller zeroes)           */
  int32_t in_dim, in_ld;     /* refnet input width (90 / 102) and row length of X0 (% 32 == 0, 32..256)          */
  int32_t pos_pe, view_pe, ref_pe;   /* frequencies of the three encodings (5, 1, 5), each <= 10                 */
  int32_t k0_dim;            /* feature channels that enter the refnet row (0 = 12).  The k0 GRID always has 12
                                channels-last channels; a model with fewer (geometry_searching: 6) keeps the rest at 0:
                                their row columns do not exist, their gradient is exactly 0                          */
  int32_t width;             /* refnet hidden width (192 / 128; 0 = 256): the GEMMs run N / K = width rounded up to 64
                                instead of the storage's 256 (tcgen05 time scales with N: 0.56x / 0.25x of the MMA work) */
  int32_t pad_;
} FgsCoarse;
int64_t fgs_coarse_wsplit_floats(int in_ld);
int fgs_coarse_forward(const FgsFineParams* p, const FgsCoarse* c, fgs_stream_t stream);
int fgs_coarse_backward(const FgsFineParams* p, const FgsFineLoss* l, const FgsCoarse* c, fgs_stream_t stream);
/* G[3][X][Y][Z] = nerf.neus_sdf_gradient('interpolate') (model/nerf.py:485-494): central differences / 2 / voxel_size
 * on interior planes, 0 on the faces; and its transpose: g_sdf += dG/dsdf^T g_G. */
int fgs_sdf_gradient_grid(const float* sdf, float* G, int X, int Y, int Z, float voxel_size, fgs_stream_t stream);
int fgs_sdf_gradient_grid_bwd(const float* g_G, float* g_sdf, int X, int Y, int Z, float voxel_size, fgs_stream_t stream);
/* transpose of fgs_conv3d_replicate (k^3 weights, replicate padding): g_in += conv^T g_out.  weights: HOST [k][k][k]. */
int fgs_conv3d_replicate_bwd(const float* g_out, float* g_in, int X, int Y, int Z, const float* weights, int ksize,
                             fgs_stream_t stream);
/* nerf.density_total_variation(sdf_tv) through autograd (ori_tv, model/nerf.py:430-447, total_variation :1212-1221):
 * loss = weight * (sum over the three axes of |sdf[v+e] - sdf[v]| with both voxels in `mask`) / 3 / mask_count;
 * g_sdf += d loss / d sdf; *loss_out += loss.  `mask` (the non-empty mask, [X][Y][Z] bytes) is required: the
 * reference's unmasked branch divides by sdf.sum(), which no config reaches (_set_nonempty_mask always runs). */
int fgs_l1_tv_add_grad(const float* sdf, const uint8_t* mask, float* g_sdf, int X, int Y, int Z, double weight,
                       double mask_count, double* loss_out, fgs_stream_t stream);
/* The unmasked branch of the same function (geometry_searching stage: no non-empty mask), as the reference writes it:
 * loss = weight * S / 3 / sum(sdf), S = sum over the three axes of |sdf[v+e] - sdf[v]| -- the divisor is the SUM OF THE
 * GRID (model/nerf.py:1221), so the gradient has a second, constant term -weight S / (3 sum^2).  Two launches: S and
 * sum into scratch2 (2 device doubles, cleared here), then the gradient. */
int fgs_l1_tv_unmasked_add_grad(const float* sdf, float* g_sdf, int X, int Y, int Z, double weight, double* scratch2,
                                double* loss_out, fgs_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* FGS_B200_FINE_H */
