/*
 * fgs_b200.h -- C-ABI of the B200-native (sm_100a) FGS-NeRF per-ray hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  Every entry
 * point names the reference interface it replaces (paths relative to the reference root,
 * yosugahhh/FGS-nerf).  All pointers are DEVICE pointers unless a parameter is documented
 * as host; all tensors are dense/contiguous fp32 unless stated; `stream` is a cudaStream_t
 * passed as void*.  Every function only ENQUEUES work on `stream` (no host sync) and
 * returns 0 on success, a cudaError_t (>0) on a CUDA failure, or a negative FGS_E_* code
 * for an argument error; fgs_last_error() returns a static description (thread-local).
 *
 * There is no CPU fallback behind any of these symbols.
 */
#ifndef FGS_B200_H
#define FGS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FGS_B200_ABI_VERSION 1
#define FGS_E_BADARG (-1)
#define FGS_E_NODEVICE (-2)

typedef void* fgs_stream_t;

int fgs_abi_version(void);
const char* fgs_last_error(void);
/* number of kernels this library has launched since load / since fgs_reset_launch_count()
 * (bench.py reports it as gpu_launches) */
int64_t fgs_launch_count(void);
void fgs_reset_launch_count(void);
/* One process drives one GPU: select it once (the library links its own CUDA runtime, so the
 * host must tell it which device the pointers and streams belong to). */
int fgs_set_device(int device);
int fgs_device_sm_count(int device, int* sm_count, int* cc_major, int* cc_minor);

/* Per-launch device timing for bench.py / profiles: between fgs_prof_begin(stream) and
 * fgs_prof_end() every launch of this library records a CUDA event on its stream.
 * fgs_prof_end() synchronises and returns the number of timed intervals (or -1);
 * fgs_prof_get(i, &name, &ms) gives interval i = the launch named `name` (time since the
 * previous launch finished, so stream-serialised kernels are timed one by one). */
int fgs_prof_begin(fgs_stream_t stream);
int fgs_prof_end(void);
int fgs_prof_get(int i, const char** name, float* ms);
/* record a host-enqueued, non-library operation in the same timing stream: what = 1 all-reduce,
 * 2 H2D copy, 3 D2H copy, 0 other */
int fgs_prof_mark(int what, fgs_stream_t stream);
/* cudaMemsetAsync(ptr, 0, bytes) on `stream` (grad buffers are cleared through the library so the
 * clears show up in the launch timing above). */
int fgs_memset_zero(void* ptr, int64_t bytes, fgs_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Ray sampling  (reference: model/cuda/render_utils.cpp:50-85, render_utils_kernel.cu:12-242)
 * ------------------------------------------------------------------------------------ */

/* render_utils_cuda.infer_t_minmax (render_utils.cpp:50-58; kernel.cu:12-35,82-104) */
int fgs_infer_t_minmax(const float* rays_o, const float* rays_d, const float* xyz_min,
                       const float* xyz_max, float near_, float far_, int n_rays,
                       float* t_min, float* t_max, fgs_stream_t stream);

/* render_utils_cuda.infer_n_samples (render_utils.cpp:60-65; kernel.cu:38-55,106-121) */
int fgs_infer_n_samples(const float* rays_d, const float* t_min, const float* t_max,
                        float stepdist, int n_rays, int64_t* n_samples, fgs_stream_t stream);

/* render_utils_cuda.infer_ray_start_dir (render_utils.cpp:67-72; kernel.cu:58-79,123-139) */
int fgs_infer_ray_start_dir(const float* rays_o, const float* rays_d, const float* t_min,
                            int n_rays, float* rays_start, float* rays_dir, fgs_stream_t stream);

/* render_utils_cuda.sample_pts_on_rays (render_utils.cpp:74-85; kernel.cu:196-242), split in
 * a count launcher and a fill launcher because the output length is data dependent.
 *
 * count: ONE kernel = infer_t_minmax + infer_n_samples + the cumsum of kernel.cu:211 and the
 *        sum of kernel.cu:212.  Writes t_min,t_max [n_rays], n_steps [n_rays] (int64),
 *        ray_offset [n_rays+1] (int64 exclusive prefix; ray_offset[n_rays] = total_len).
 *        `scratch` is >= fgs_sample_scratch_bytes(n_rays) bytes of device memory that must
 *        be zero before the first call; the kernel leaves it zeroed again.
 * fill : ONE kernel = __set_1_at_ray_seg_start + cumsum_ + __set_step_id +
 *        infer_ray_start_dir + sample_pts_on_rays_cuda_kernel (kernel.cu:144-194,213-240).
 *        total_len is the host copy of ray_offset[n_rays]. */
int64_t fgs_sample_scratch_bytes(int n_rays);
int fgs_sample_pts_count(const float* rays_o, const float* rays_d, const float* xyz_min,
                         const float* xyz_max, float near_, float far_, float stepdist,
                         int n_rays, float* t_min, float* t_max, int64_t* n_steps,
                         int64_t* ray_offset, void* scratch, fgs_stream_t stream);
int fgs_sample_pts_fill(const float* rays_o, const float* rays_d, const float* xyz_min,
                        const float* xyz_max, const float* t_min, const int64_t* ray_offset,
                        float stepdist, int n_rays, int64_t total_len, float* ray_pts,
                        uint8_t* mask_outbbox, int64_t* ray_id, int64_t* step_id,
                        fgs_stream_t stream);

/* render_utils_cuda.sample_ndc_pts_on_rays (render_utils.cpp:87-97; kernel.cu:245-293) */
int fgs_sample_ndc_pts_on_rays(const float* rays_o, const float* rays_d, const float* xyz_min,
                               const float* xyz_max, int n_samples, int n_rays, float* ray_pts,
                               uint8_t* mask_outbbox, fgs_stream_t stream);

/* render_utils_cuda.sample_bg_pts_on_rays (render_utils.cpp:99-107; kernel.cu:301-360) */
int fgs_sample_bg_pts_on_rays(const float* rays_o, const float* rays_d, const float* t_max,
                              float bg_preserve, int n_samples, int n_rays, float* ray_pts,
                              fgs_stream_t stream);

/* render_utils_cuda.maskcache_lookup (render_utils.cpp:109-118; kernel.cu:374-424).
 * world: bool [sz_i,sz_j,sz_k]; out: bool [n_pts] (every element written, 0 if outside). */
int fgs_maskcache_lookup(const uint8_t* world, const float* xyz, const float* xyz2ijk_scale,
                         const float* xyz2ijk_shift, int sz_i, int sz_j, int sz_k,
                         int64_t n_pts, uint8_t* out, fgs_stream_t stream);

/* render_utils_cuda.raw2alpha / raw2alpha_nonuni and their backwards
 * (render_utils.cpp:120-140; kernel.cu:431-574).  interval_pts may be NULL (uniform). */
int fgs_raw2alpha(const float* density, float shift, float interval, const float* interval_pts,
                  int64_t n_pts, float* exp_d, float* alpha, fgs_stream_t stream);
int fgs_raw2alpha_backward(const float* exp_d, const float* grad_back, float interval,
                           const float* interval_pts, int64_t n_pts, float* grad,
                           fgs_stream_t stream);

/* render_utils_cuda.alpha2weight (render_utils.cpp:142-148; kernel.cu:577-651): segmented
 * transmittance scan keyed on the sorted ray_id, early stop at T<1e-3.  All five outputs are
 * fully written (weight/T tails keep 0/1, empty rays keep alphainv_last=1, i_start=i_end=0). */
int fgs_alpha2weight(const float* alpha, const int64_t* ray_id, int64_t n_pts, int n_rays,
                     float* weight, float* T, float* alphainv_last, int64_t* i_start,
                     int64_t* i_end, fgs_stream_t stream);

/* render_utils_cuda.alpha2weight_backward (render_utils.cpp:150-167; kernel.cu:654-707) */
int fgs_alpha2weight_backward(const float* alpha, const float* weight, const float* T,
                              const float* alphainv_last, const int64_t* i_start,
                              const int64_t* i_end, int n_rays, int64_t n_pts,
                              const float* grad_weights, const float* grad_last, float* grad,
                              fgs_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Optimizer  (reference: model/cuda/adam_upd.cpp:36-86, adam_upd_kernel.cu:8-132)
 * mode 0 = adam_upd, 1 = masked_adam_upd (skip grad==0), 2 = adam_upd_with_perlr.
 * ------------------------------------------------------------------------------------ */
int fgs_adam_upd(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                 const float* perlr, int64_t n, int step, float beta1, float beta2, float lr,
                 float eps, int mode, fgs_stream_t stream);

/* Multi-tensor dense Adam: one launch for up to FGS_MAX_MT tensors (the 16 MLP tensors the
 * reference updates with 16 launches, model/adam.py:190-221).  All arrays are HOST arrays. */
#define FGS_MAX_MT 32
int fgs_adam_upd_multi(int n_tensors, float* const* params, const float* const* grads,
                       float* const* exp_avgs, float* const* exp_avg_sqs, const int64_t* numels,
                       const int* steps, const float* lrs, float beta1, float beta2, float eps,
                       fgs_stream_t stream);
/* The same, clearing every gradient element it consumes (the fused step's MLP gradients are accumulated with atomics:
 * this replaces the memset before the next step's weight-gradient kernel). */
int fgs_adam_upd_multi_zero_grad(int n_tensors, float* const* params, float* const* grads,
                                 float* const* exp_avgs, float* const* exp_avg_sqs, const int64_t* numels,
                                 const int* steps, const float* lrs, float beta1, float beta2, float eps,
                                 fgs_stream_t stream);

/* Fused-step variants (same arithmetic, less traffic):
 *  fgs_adam_upd_zero_grad : dense adam_upd that also clears the consumed gradient (no memset pass);
 *  fgs_brick_list         : compacts the touched-brick flags that fgs_fine_backward leaves behind (one
 *      byte per 4x4x4-voxel brick of k0, [ceil(X/4)*ceil(Y/4)*ceil(Z/4)]) into brick_list
 *      (fgs_brick_list_ints() ints: count, then brick ids in ascending order) and clears the flags;
 *  fgs_brick_copy         : gathers (pack != 0) / scatters the listed gradient bricks to / from a compact
 *      [count][64 * channels] buffer -- the payload of the sparse multi-GPU exchange; channels 12 (k0, channels-last,
 *      768 floats per brick) or 1 (the SDF gradient, 64 floats per brick);
 *  fgs_adam_upd_bricks    : masked_adam_upd of a channels-last [X][Y][Z][12] grid restricted to the
 *      listed bricks, one warp per brick; clears the gradients it consumed.  Identical p, m, v to
 *      masked_adam_upd provided every voxel with grad != 0 lies in a listed brick. */
int fgs_adam_upd_zero_grad(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                           int step, float beta1, float beta2, float lr, float eps,
                           fgs_stream_t stream);
int64_t fgs_brick_list_ints(int X, int Y, int Z);
/* Voxel-granular versions for the data-parallel exchange (only ~23 % of the voxels of a touched brick carry a
 * gradient): voxel_flags has one byte per voxel of the [X][Y][Z][12] grid (FgsFineParams::k0_vox_flags, set by the
 * backward scatter); fgs_voxel_list writes count + ascending voxel ids into voxel_list (fgs_voxel_list_ints(V) ints,
 * the tail is scratch) with two launches over all SMs and clears the flags; fgs_voxel_copy gathers (pack != 0) /
 * scatters the 12 channels of the first min(count, max_voxels) listed voxels to / from compact [count][12]. */
int64_t fgs_voxel_list_ints(int64_t n_voxels);
/* Bit-packed variant for the multi-rank union: fgs_voxel_bits_pack turns the byte flags into one word per 32 voxels
 * (and clears them); after an all-gather of the ranks' masks into bits_all [world][n_voxels / 32],
 * fgs_voxel_bits_union_list ORs them (union left in slice 0), sets the touched-brick byte flags of the union (may be
 * NULL) and writes count + ascending voxel ids into voxel_list (fgs_voxel_list_ints ints).  Needs Z % 4 == 0 and
 * X*Y*Z % 32 == 0. */
int fgs_voxel_bits_pack(uint8_t* voxel_flags, uint32_t* bits, int64_t n_voxels, fgs_stream_t stream);
int fgs_voxel_bits_union_list(uint32_t* bits_all, int world, int X, int Y, int Z, uint8_t* brick_flags,
                              int32_t* voxel_list, fgs_stream_t stream);
int fgs_voxel_list(uint8_t* voxel_flags, int64_t n_voxels, int32_t* voxel_list, fgs_stream_t stream);
int fgs_voxel_copy(float* grid, float* compact, const int32_t* voxel_list, int max_voxels, int channels, int pack,
                   fgs_stream_t stream);
int fgs_brick_list(uint8_t* brick_flags, int X, int Y, int Z, int32_t* brick_list, fgs_stream_t stream);
int fgs_brick_copy(float* grid, float* compact, const int32_t* brick_list, int max_bricks, int X, int Y,
                   int Z, int channels, int pack, fgs_stream_t stream);
int fgs_adam_upd_bricks(float* param, float* grad, float* exp_avg, float* exp_avg_sq,
                        const int32_t* brick_list, int X, int Y, int Z, int channels, int step,
                        float beta1, float beta2, float lr, float eps, fgs_stream_t stream);
/* The same masked Adam for ONE GPU, where nobody needs the ordered list (it exists to pack / unpack the multi-GPU
 * exchange): the flags are compacted over all SMs in arrival order into brick_ulist (fgs_brick_list_ints() + 1 ints,
 * zero-initialised once: two alternating counters keyed on `step & 1`, then ids) and cleared, then one warp per listed
 * brick.  `step` must advance by one per call. */
int fgs_adam_upd_brick_flags(float* param, float* grad, float* exp_avg, float* exp_avg_sq, uint8_t* brick_flags,
                             int32_t* brick_ulist, int X, int Y, int Z, int channels, int step, float beta1,
                             float beta2, float lr, float eps, fgs_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Total variation  (reference: model/cuda/total_variation.cpp:16-32, _kernel.cu:13-133)
 * param/grad: [1,C,X,Y,Z]; channels_last=0 is the reference's contiguous NCDHW layout,
 * channels_last=1 is torch.channels_last_3d storage (C innermost) of the same logical tensor.
 * mask (may be NULL): float, same layout -> total_variation_add_grad_new.
 * ------------------------------------------------------------------------------------ */
int fgs_total_variation_add_grad(const float* param, float* grad, const float* mask, float wx,
                                 float wy, float wz, int C, int sz_i, int sz_j, int sz_k,
                                 int dense_mode, int channels_last, fgs_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* FGS_B200_H */
