/*
 * fgs_b200_train.h -- C-ABI of the pieces AROUND the per-ray step that SURVEY.md 8(f) ranks next:
 * the training loop's batch gather (f2) and the progressive grid scaling between stages (f3).
 * Same conventions as fgs_b200.h: device pointers, dense fp32, enqueue-only, 0 on success.
 */
#ifndef FGS_B200_TRAIN_H
#define FGS_B200_TRAIN_H

#include "fgs_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------
 * Progressive scaling  (reference: DenseGrid.scale_volume_grid, model/grid.py:101-106, called by
 * nerf.scale_volume_grid, model/nerf.py:369-381, at the pg_scale iterations, nerf_training.py:244-254)
 *
 *   dst[1,C,X2,Y2,Z2] = F.interpolate(src[1,C,X,Y,Z], size=(X2,Y2,Z2), mode='trilinear', align_corners=True)
 *
 * ATen's upsample_trilinear3d arithmetic (float scale = (in-1)/(out-1), 0 when out == 1; source index
 * = scale * dst index; truncation; lambda1 = frac, lambda0 = 1 - lambda1; nested w -> h -> d
 * blend).  channels_last = 0: both tensors contiguous NCDHW (the reference layout, the SDF grid);
 * channels_last = 1: both in torch.channels_last_3d storage, C innermost (the engine's k0 layout;
 * needs C % 4 == 0).  One streaming pass: every output element is written once with 128-bit stores,
 * the (smaller) source is read through L1 / L2.
 * ------------------------------------------------------------------------------------ */
int fgs_grid_resize_trilinear(const float* src, int C, int X, int Y, int Z, float* dst, int X2, int Y2,
                              int Z2, int channels_last, fgs_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Batch gather  (reference: the `flatten` / `in_maskcache` ray sampler of the training loop,
 * model/nerf_training.py:253-258:  target = rgb_tr[sel_i]; rays_o = rays_o_tr[sel_i]; ... )
 *
 * sel [n] int64 indices into the flattened training rays [n_total,3]; the four selections are one
 * launch.  Indices outside [0, n_total) are an error in the reference (device assert); here they set
 * *bad (device int32, may be NULL) to 1 and the row is zero-filled.
 * ------------------------------------------------------------------------------------ */
int fgs_gather_rays(const float* rays_o_tr, const float* rays_d_tr, const float* viewdirs_tr,
                    const float* rgb_tr, int64_t n_total, const int64_t* sel, int n, float* rays_o,
                    float* rays_d, float* viewdirs, float* target, int32_t* bad, fgs_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Gaussian smoothing of the SDF volume  (reference: nerf._gaussian_3dconv, model/nerf.py:260-272 =
 * nn.Conv3d(1, 1, ksize, stride=1, padding=ksize//2, padding_mode='replicate'), zero bias; used at the stage
 * hand-off by nerf.init_sdf_from_sdf, model/nerf.py:280-296, called from model/nerf_training.py:128)
 *
 *   dst[x][y][z] = ( sum_{a,b,c} w[a][b][c] * (src[clamp(x+a-h)][clamp(y+b-h)][clamp(z+c-h)] / pre_div) ) / post_div
 *
 * weights: HOST pointer to ksize^3 floats in [a][b][c] order (the caller builds them exactly as the reference
 * does: float64 exp, cast to fp32, divided by their fp32 sum); ksize in {1,3,5,7}; pre_div / post_div fold
 * init_sdf_from_sdf's `m(sdf0 / reduce) / reduce` into the pass (pass 1 to skip), evaluated as torch's CUDA
 * tensor / python-scalar division is: x * (1.0f / divisor).  One shared-memory tiled pass:
 * 8 bytes per voxel of compulsory traffic.  The accumulation order ((b, c) outer, a inner, FMA) is this kernel's own --
 * cuDNN's is unspecified -- so parity with the reference is to fp32 rounding (tests: 2e-6 of max|dst|).
 * ------------------------------------------------------------------------------------ */
int fgs_conv3d_replicate(const float* src, float* dst, int X, int Y, int Z, const float* weights, int ksize,
                         float pre_div, float post_div, fgs_stream_t stream);

/* ------------------------------------------------------------------------------------
 * SDF field query for mesh extraction  (reference: extract_fields, model/extract_geometry.py:5-19, with the
 * query_func of nerf.extract_geometry, model/nerf.py:1157-1170: grid_sampler(pts, -sdf_grid))
 *
 *   out[i][j][k] = sign * trilinear(grid [X][Y][Z], (xs[i], ys[j], zs[k]))      (zero padding outside the bbox)
 *
 * xs / ys / zs are the three coordinate arrays (device; the reference builds them with torch.linspace), so the
 * lattice points never exist as a [M,3] tensor: 4 bytes written per point instead of 12 written + 12 read + 4
 * written, and no host round trip per 64^3 block.  xyz_min / xyz_max are HOST pointers to 3 floats (the grid's
 * bbox).  sign = -1 gives the reference's field (it samples the negated grid).  Same arithmetic as
 * fgs_grid_sample_fwd (ATen grid_sampler_3d), so values are bit-identical to the reference's on CUDA.
 * ------------------------------------------------------------------------------------ */
int fgs_grid_sample_lattice(const float* grid, int X, int Y, int Z, const float* xyz_min, const float* xyz_max,
                            const float* xs, int nx, const float* ys, int ny, const float* zs, int nz, float sign,
                            float* out, fgs_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* FGS_B200_TRAIN_H */
