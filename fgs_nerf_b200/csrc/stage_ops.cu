// Between-step operators of the training loop: progressive grid scaling and the batch gather.
// Both are single-pass streaming kernels: HBM-bound, 128-bit stores, grids sized in multiples of the SM count.
#include "common.cuh"
#include "fine_common.cuh"
#include "../../include/fgs_b200_train.h"

namespace {

constexpr int kBlock = 256;

// ATen area_pixel_compute_scale<float>(in, out, align_corners=true): (float)(in - 1) / (out - 1), 0 if out <= 1
inline float resize_scale(int in, int out) { return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f; }

struct Axis {
  int i0, ip;        // lower source index, step to the upper one (0 on the last plane)
  float l0, l1;      // weights of the lower / upper sample
};

// upsample_trilinear3d (align_corners): src = scale * dst; i0 = (int)src; l1 = src - i0; l0 = 1 - l1
__device__ __forceinline__ Axis axis_of(int dst, float scale, int in) {
  Axis a;
  const float r = __fmul_rn(scale, (float)dst);
  a.i0 = (int)r;
  a.ip = (a.i0 < in - 1) ? 1 : 0;
  a.l1 = __fsub_rn(r, (float)a.i0);
  a.l0 = __fsub_rn(1.f, a.l1);
  return a;
}

// val = d0 * (h0 * (w0 * v000 + w1 * v001) + h1 * (w0 * v010 + w1 * v011)) + d1 * (same on the upper plane)
__device__ __forceinline__ float blend(const Axis& d, const Axis& h, const Axis& w, float v000, float v001, float v010,
                                       float v011, float v100, float v101, float v110, float v111) {
  const float lo = h.l0 * (w.l0 * v000 + w.l1 * v001) + h.l1 * (w.l0 * v010 + w.l1 * v011);
  const float hi = h.l0 * (w.l0 * v100 + w.l1 * v101) + h.l1 * (w.l0 * v110 + w.l1 * v111);
  return d.l0 * lo + d.l1 * hi;
}

// Both resize kernels walk the output row by row so that no thread divides: a CTA owns a strip of output rows
// (fixed x2, a few consecutive y2), decodes it once from blockIdx, and its threads stride along z.  (The first version
// decoded a flat 64-bit item index per thread -- three emulated 64-bit divisions per 16 bytes stored -- and ran at
// 27 % of the HBM roofline.)
constexpr int RS_TZ = 64, RS_TY = 4;   // thread layout of a CTA: 64 along z x 4 rows
static_assert(RS_TZ * RS_TY == kBlock, "resize thread layout");
constexpr int RS_YB = 16;              // output rows (y2) per CTA

// NCDHW: one thread = VEC consecutive outputs along Z2 of one (c, x2, y2) row.
template <int VEC>
__global__ void __launch_bounds__(kBlock) resize_ncdhw_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                              int C, int X, int Y, int Z, int X2, int Y2, int Z2,
                                                              float sx, float sy, float sz, int y_strips,
                                                              long long n_ctas) {
  const int Zq = Z2 / VEC;
  const int tz = threadIdx.x % RS_TZ, ty = threadIdx.x / RS_TZ;
  for (long long cta = blockIdx.x; cta < n_ctas; cta += gridDim.x) {
    const int ys = (int)(cta % y_strips);
    const long long r = cta / y_strips;
    const int x2 = (int)(r % X2), c = (int)(r / X2);
    const Axis ax = axis_of(x2, sx, X);
    const float* plane0 = src + ((long long)c * X + ax.i0) * Y * Z;
    const float* plane1 = plane0 + (long long)ax.ip * Y * Z;
    const int y_end = min(Y2, (ys + 1) * RS_YB);
    for (int y2 = ys * RS_YB + ty; y2 < y_end; y2 += RS_TY) {
      const Axis ay = axis_of(y2, sy, Y);
      const float* p00 = plane0 + (long long)ay.i0 * Z;
      const float* p01 = p00 + (long long)ay.ip * Z;
      const float* p10 = plane1 + (long long)ay.i0 * Z;
      const float* p11 = p10 + (long long)ay.ip * Z;
      float* orow = dst + (((long long)c * X2 + x2) * Y2 + y2) * Z2;
      for (int zq = tz; zq < Zq; zq += RS_TZ) {
        float out[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          const Axis az = axis_of(zq * VEC + e, sz, Z);
          const int z0 = az.i0, z1 = az.i0 + az.ip;
          out[e] = blend(ax, ay, az, __ldg(p00 + z0), __ldg(p00 + z1), __ldg(p01 + z0), __ldg(p01 + z1),
                         __ldg(p10 + z0), __ldg(p10 + z1), __ldg(p11 + z0), __ldg(p11 + z1));
        }
        if (VEC == 4) {
          st4(orow + zq * 4, make_float4(out[0], out[1], out[2], out[3]));
        } else {
#pragma unroll
          for (int e = 0; e < VEC; ++e) orow[zq * VEC + e] = out[e];
        }
      }
    }
  }
}

// channels-last ([X][Y][Z][C]): one thread = all C channels of one output voxel, four at a time; the 8 corners are
// 128-bit loads, the axis weights are shared by the C/4 quads.
constexpr int RC_YB = 4;               // output rows (y2) per CTA
__global__ void __launch_bounds__(kBlock) resize_cl_kernel(const float* __restrict__ src, float* __restrict__ dst, int C,
                                                           int X, int Y, int Z, int X2, int Y2, int Z2, float sx,
                                                           float sy, float sz, int y_strips, long long n_ctas) {
  const int Cq = C >> 2;
  const long long sZ = C, sY = (long long)Z * C, sX = (long long)Y * Z * C;
  for (long long cta = blockIdx.x; cta < n_ctas; cta += gridDim.x) {
    const int ys = (int)(cta % y_strips), x2 = (int)(cta / y_strips);
    const Axis ax = axis_of(x2, sx, X);
    const int y_end = min(Y2, (ys + 1) * RC_YB);
    for (int y2 = ys * RC_YB; y2 < y_end; ++y2) {
      const Axis ay = axis_of(y2, sy, Y);
      const float* row = src + ax.i0 * sX + ay.i0 * sY;
      const long long dx = ax.ip * sX, dy = ay.ip * sY;
      float* orow = dst + ((long long)x2 * Y2 + y2) * Z2 * C;
      for (int z2 = threadIdx.x; z2 < Z2; z2 += kBlock) {
        const Axis az = axis_of(z2, sz, Z);
        const float* p = row + az.i0 * sZ;
        const long long dz = az.ip * sZ;
        for (int cq = 0; cq < Cq; ++cq, p += 4) {
          const float4 v000 = ld4(p), v001 = ld4(p + dz), v010 = ld4(p + dy), v011 = ld4(p + dy + dz);
          const float4 v100 = ld4(p + dx), v101 = ld4(p + dx + dz), v110 = ld4(p + dx + dy), v111 = ld4(p + dx + dy + dz);
          float4 o;
          o.x = blend(ax, ay, az, v000.x, v001.x, v010.x, v011.x, v100.x, v101.x, v110.x, v111.x);
          o.y = blend(ax, ay, az, v000.y, v001.y, v010.y, v011.y, v100.y, v101.y, v110.y, v111.y);
          o.z = blend(ax, ay, az, v000.z, v001.z, v010.z, v011.z, v100.z, v101.z, v110.z, v111.z);
          o.w = blend(ax, ay, az, v000.w, v001.w, v010.w, v011.w, v100.w, v101.w, v110.w, v111.w);
          st4(orow + (long long)z2 * C + cq * 4, o);
        }
      }
    }
  }
}

// one thread = one (ray, array) pair: 3 floats in, 3 floats out
__global__ void __launch_bounds__(kBlock) gather_rays_kernel(const float* __restrict__ o_tr, const float* __restrict__ d_tr,
                                                             const float* __restrict__ v_tr, const float* __restrict__ c_tr,
                                                             long long n_total, const int64_t* __restrict__ sel, int n,
                                                             float* __restrict__ o, float* __restrict__ d,
                                                             float* __restrict__ v, float* __restrict__ c,
                                                             int32_t* __restrict__ bad) {
  const int t = blockIdx.x * kBlock + threadIdx.x;
  if (t >= 4 * n) return;
  const int which = t / n, i = t - which * n;      // array-major: a warp reads neighbouring sel entries
  const float* in = which == 0 ? o_tr : which == 1 ? d_tr : which == 2 ? v_tr : c_tr;
  float* out = which == 0 ? o : which == 1 ? d : which == 2 ? v : c;
  const long long s = sel[i];
  float a = 0.f, b = 0.f, e = 0.f;
  if (s >= 0 && s < n_total) {
    a = __ldg(in + 3 * s);
    b = __ldg(in + 3 * s + 1);
    e = __ldg(in + 3 * s + 2);
  } else if (bad) {
    *bad = 1;
  }
  out[3 * (long long)i] = a;
  out[3 * (long long)i + 1] = b;
  out[3 * (long long)i + 2] = e;
}

// SDF field on a regular lattice (extract_fields): one thread per lattice point, z fastest (coalesced stores);
// the coordinates come from the three 1-D arrays the caller built with torch.linspace, the sample is ATen's
// grid_sampler_3d arithmetic (ray_math.cuh), so the values equal the reference's query_func bit for bit.
__global__ void __launch_bounds__(kBlock) lattice_sample_kernel(const float* __restrict__ grid, fine::GridGeo g,
                                                                const float* __restrict__ xs, int nx,
                                                                const float* __restrict__ ys, int ny,
                                                                const float* __restrict__ zs, int nz, float sign,
                                                                float* __restrict__ out, int y_strips, long long n_ctas) {
  // a CTA owns 4 lattice rows (fixed i, 4 consecutive j), threads stride along k: no per-thread division, the x / y
  // continuous indices are computed once per row
  for (long long cta = blockIdx.x; cta < n_ctas; cta += gridDim.x) {
    const int js = (int)(cta % y_strips), i = (int)(cta / y_strips);
    const float id = aten_unnormalize(ref_normalize(__ldg(xs + i), g.mn[0], g.ext[0]), g.X);
    const int j_end = min(ny, (js + 1) * 4);
    for (int j = js * 4; j < j_end; ++j) {
      const float ih = aten_unnormalize(ref_normalize(__ldg(ys + j), g.mn[1], g.ext[1]), g.Y);
      float* orow = out + ((long long)i * ny + j) * nz;
      for (int k = threadIdx.x; k < nz; k += kBlock) {
        const float iw = aten_unnormalize(ref_normalize(__ldg(zs + k), g.mn[2], g.ext[2]), g.Z);
        const Tri t = tri_setup(iw, ih, id);
        orow[k] = __fmul_rn(sign, tri_sample1(grid, t, g.X, g.Y, g.Z));
      }
    }
  }
}

// Dense k^3 convolution of a single-channel volume with replicate padding (nn.Conv3d(1, 1, k, padding=k//2,
// padding_mode='replicate'), the reference's Gaussian smoother).  A CTA computes a 4 x 8 x 32 block of outputs from a
// shared-memory tile with its halo (each input voxel is read from HBM ~1.5-2x instead of k^3 times); the k^3 weights
// travel as a kernel parameter (constant bank, uniform index).  pre_div / post_div fold the reference's
// `m(sdf0 / reduce) / reduce` into the same pass the way torch's CUDA kernel divides a tensor by a python scalar:
// one fp32 reciprocal on the host, a multiplication per element (skipped when the divisor is 1).
constexpr int kConvMaxK = 7;
constexpr int CTX = 4, CTY = 8, CTZ = 32;
struct ConvWeights { float w[kConvMaxK * kConvMaxK * kConvMaxK]; };

template <int K>
__global__ void __launch_bounds__(kBlock) conv3d_replicate_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                                 int X, int Y, int Z, ConvWeights cw, float pre_inv,
                                                                 float post_inv, int tiles_y, int tiles_z) {
  constexpr int H = K / 2;
  constexpr int SX = CTX + 2 * H, SY = CTY + 2 * H, SZ = CTZ + 2 * H;
  __shared__ float tile[SX * SY * SZ];
  const int tz = blockIdx.x % tiles_z, ty = (blockIdx.x / tiles_z) % tiles_y, tx = blockIdx.x / (tiles_z * tiles_y);
  const int x0 = tx * CTX, y0 = ty * CTY, z0 = tz * CTZ;
  for (int i = threadIdx.x; i < SX * SY * SZ; i += kBlock) {
    const int lz = i % SZ, ly = (i / SZ) % SY, lx = i / (SZ * SY);
    const int gx = min(max(x0 + lx - H, 0), X - 1), gy = min(max(y0 + ly - H, 0), Y - 1), gz = min(max(z0 + lz - H, 0), Z - 1);
    float v = __ldg(src + ((long long)gx * Y + gy) * Z + gz);
    if (pre_inv != 1.f) v = __fmul_rn(v, pre_inv);
    tile[i] = v;
  }
  __syncthreads();
  // 32 x 8 threads; a thread keeps the CTX outputs of its (y, z) column in registers: for each (b, c) it reads the
  // CTX + K - 1 tile values of the column once and feeds CTX x K FMAs (2.5x fewer shared-memory reads than one read
  // per tap at K = 5 -- the first version was bound by them: 125 LDS per output)
  const int lz = threadIdx.x % CTZ, ly = threadIdx.x / CTZ;
  float acc[CTX];
#pragma unroll
  for (int lx = 0; lx < CTX; ++lx) acc[lx] = 0.f;
  // accumulation order per output: (b, c) outer, a inner -- fixed, documented in the header
#pragma unroll 1
  for (int b = 0; b < K; ++b) {
#pragma unroll
    for (int c = 0; c < K; ++c) {
      float col[CTX + K - 1];
#pragma unroll
      for (int i = 0; i < CTX + K - 1; ++i) col[i] = tile[(i * SY + (ly + b)) * SZ + lz + c];
#pragma unroll
      for (int a = 0; a < K; ++a) {
        const float w = cw.w[(a * K + b) * K + c];
#pragma unroll
        for (int lx = 0; lx < CTX; ++lx) acc[lx] = __fmaf_rn(w, col[lx + a], acc[lx]);
      }
    }
  }
#pragma unroll
  for (int lx = 0; lx < CTX; ++lx) {
    const int gx = x0 + lx, gy = y0 + ly, gz = z0 + lz;
    if (gx < X && gy < Y && gz < Z) {
      float v = acc[lx];
      if (post_inv != 1.f) v = __fmul_rn(v, post_inv);
      dst[((long long)gx * Y + gy) * Z + gz] = v;
    }
  }
}

template <int K>
int launch_conv(const float* src, float* dst, int X, int Y, int Z, const float* w, float pre_div, float post_div,
                cudaStream_t st) {
  ConvWeights cw;
  for (int i = 0; i < K * K * K; ++i) cw.w[i] = w[i];
  const int tx = fgs::cdiv(X, CTX), ty = fgs::cdiv(Y, CTY), tz = fgs::cdiv(Z, CTZ);
  const long long blocks = (long long)tx * ty * tz;
  if (blocks > 0x7fffffffLL) return fgs::fail_arg("conv3d_replicate: volume too large");
  // ATen's div_true_kernel_cuda with a CPU-scalar divisor: inv_b = 1.0f / b, then a * inv_b
  const float pre_inv = pre_div == 1.f ? 1.f : 1.0f / pre_div, post_inv = post_div == 1.f ? 1.f : 1.0f / post_div;
  conv3d_replicate_kernel<K><<<(unsigned)blocks, kBlock, 0, st>>>(src, dst, X, Y, Z, cw, pre_inv, post_inv, ty, tz);
  return fgs::check_launch("conv3d_replicate", st);
}

// Transpose of the convolution above (its backward with respect to the input volume).  Replicate padding = the same
// correlation over a padded copy P[p] = src[clamp(p)], p in [-H, n + H): so g_P[p] = sum_a w[a] g_out[p - a + H] (g_out read
// as ZERO outside the volume) = the forward kernel's loop with flipped weights over the extended domain, and
// g_src[clamp(p)] += g_P[p].  Interior voxels receive exactly one add; only the faces fold several p onto one voxel.
template <int K>
__global__ void __launch_bounds__(kBlock) conv3d_replicate_bwd_kernel(const float* __restrict__ g_out, float* __restrict__ g_in,
                                                                     int X, int Y, int Z, ConvWeights cw, int tiles_y,
                                                                     int tiles_z) {
  constexpr int H = K / 2;
  constexpr int SX = CTX + 2 * H, SY = CTY + 2 * H, SZ = CTZ + 2 * H;
  __shared__ float tile[SX * SY * SZ];
  const int tz = blockIdx.x % tiles_z, ty = (blockIdx.x / tiles_z) % tiles_y, tx = blockIdx.x / (tiles_z * tiles_y);
  const int x0 = tx * CTX - H, y0 = ty * CTY - H, z0 = tz * CTZ - H;          // padded coordinates of this block
  for (int i = threadIdx.x; i < SX * SY * SZ; i += kBlock) {
    const int lz = i % SZ, ly = (i / SZ) % SY, lx = i / (SZ * SY);
    const int gx = x0 + lx - H, gy = y0 + ly - H, gz = z0 + lz - H;
    const bool in = gx >= 0 && gx < X && gy >= 0 && gy < Y && gz >= 0 && gz < Z;
    tile[i] = in ? __ldg(g_out + ((long long)gx * Y + gy) * Z + gz) : 0.f;
  }
  __syncthreads();
  const int lz = threadIdx.x % CTZ, ly = threadIdx.x / CTZ;
  float acc[CTX];
#pragma unroll
  for (int lx = 0; lx < CTX; ++lx) acc[lx] = 0.f;
#pragma unroll 1
  for (int b = 0; b < K; ++b) {
#pragma unroll
    for (int c = 0; c < K; ++c) {
      float col[CTX + K - 1];
#pragma unroll
      for (int i = 0; i < CTX + K - 1; ++i) col[i] = tile[(i * SY + (ly + b)) * SZ + lz + c];
#pragma unroll
      for (int a = 0; a < K; ++a) {
        const float w = cw.w[((K - 1 - a) * K + (K - 1 - b)) * K + (K - 1 - c)];
#pragma unroll
        for (int lx = 0; lx < CTX; ++lx) acc[lx] = __fmaf_rn(w, col[lx + a], acc[lx]);
      }
    }
  }
#pragma unroll
  for (int lx = 0; lx < CTX; ++lx) {
    const int px = x0 + lx, py = y0 + ly, pz = z0 + lz;
    if (px < X + H && py < Y + H && pz < Z + H && acc[lx] != 0.f) {
      const int gx = min(max(px, 0), X - 1), gy = min(max(py, 0), Y - 1), gz = min(max(pz, 0), Z - 1);
      atomicAdd(g_in + ((long long)gx * Y + gy) * Z + gz, acc[lx]);
    }
  }
}

template <int K>
int launch_conv_bwd(const float* g_out, float* g_in, int X, int Y, int Z, const float* w, cudaStream_t st) {
  ConvWeights cw;
  for (int i = 0; i < K * K * K; ++i) cw.w[i] = w[i];
  constexpr int H = K / 2;
  const int tx = fgs::cdiv(X + 2 * H, CTX), ty = fgs::cdiv(Y + 2 * H, CTY), tz = fgs::cdiv(Z + 2 * H, CTZ);
  const long long blocks = (long long)tx * ty * tz;
  if (blocks > 0x7fffffffLL) return fgs::fail_arg("conv3d_replicate_bwd: volume too large");
  conv3d_replicate_bwd_kernel<K><<<(unsigned)blocks, kBlock, 0, st>>>(g_out, g_in, X, Y, Z, cw, ty, tz);
  return fgs::check_launch("conv3d_replicate_bwd", st);
}

}  // namespace

extern "C" int fgs_conv3d_replicate(const float* src, float* dst, int X, int Y, int Z, const float* weights, int ksize,
                                    float pre_div, float post_div, fgs_stream_t stream) {
  FGS_REQUIRE(X > 0 && Y > 0 && Z > 0, "conv3d_replicate: sizes must be positive");
  FGS_REQUIRE(ksize == 1 || ksize == 3 || ksize == 5 || ksize == 7, "conv3d_replicate: ksize must be 1, 3, 5 or 7");
  FGS_REQUIRE(src && dst && src != dst && weights, "conv3d_replicate: src / dst must be distinct non-null buffers");
  FGS_REQUIRE(pre_div != 0.f && post_div != 0.f, "conv3d_replicate: zero divisor");
  cudaStream_t st = (cudaStream_t)stream;
  switch (ksize) {
    case 1: return launch_conv<1>(src, dst, X, Y, Z, weights, pre_div, post_div, st);
    case 3: return launch_conv<3>(src, dst, X, Y, Z, weights, pre_div, post_div, st);
    case 5: return launch_conv<5>(src, dst, X, Y, Z, weights, pre_div, post_div, st);
    default: return launch_conv<7>(src, dst, X, Y, Z, weights, pre_div, post_div, st);
  }
}

extern "C" int fgs_grid_sample_lattice(const float* grid, int X, int Y, int Z, const float* xyz_min,
                                       const float* xyz_max, const float* xs, int nx, const float* ys, int ny,
                                       const float* zs, int nz, float sign, float* out, fgs_stream_t stream) {
  FGS_REQUIRE(X > 0 && Y > 0 && Z > 0 && nx >= 0 && ny >= 0 && nz >= 0, "grid_sample_lattice: bad size");
  const long long n = (long long)nx * ny * nz;
  if (n == 0) return 0;
  FGS_REQUIRE(grid && xs && ys && zs && out && xyz_min && xyz_max, "grid_sample_lattice: null pointer");
  fine::GridGeo g;
  for (int c = 0; c < 3; ++c) {
    g.mn[c] = xyz_min[c];
    g.ext[c] = xyz_max[c] - xyz_min[c];  // fp32 subtraction, as torch evaluates (xyz_max - xyz_min)
  }
  g.X = X; g.Y = Y; g.Z = Z;
  cudaStream_t st = (cudaStream_t)stream;
  const int y_strips = fgs::cdiv(ny, 4);
  const long long n_ctas = (long long)nx * y_strips;
  const long long cap = FGS_NUM_SMS * 16;
  lattice_sample_kernel<<<(unsigned)(n_ctas < cap ? n_ctas : cap), kBlock, 0, st>>>(grid, g, xs, nx, ys, ny, zs, nz, sign,
                                                                                      out, y_strips, n_ctas);
  return fgs::check_launch("grid_sample_lattice", st);
}

extern "C" int fgs_grid_resize_trilinear(const float* src, int C, int X, int Y, int Z, float* dst, int X2, int Y2,
                                         int Z2, int channels_last, fgs_stream_t stream) {
  FGS_REQUIRE(C >= 0 && X > 0 && Y > 0 && Z > 0 && X2 > 0 && Y2 > 0 && Z2 > 0, "grid_resize: sizes must be positive");
  if (C == 0) return 0;
  FGS_REQUIRE(src && dst && src != dst, "grid_resize: src / dst must be distinct non-null buffers");
  cudaStream_t st = (cudaStream_t)stream;
  const float sx = resize_scale(X, X2), sy = resize_scale(Y, Y2), sz = resize_scale(Z, Z2);
  const int cap = FGS_NUM_SMS * 16;
  if (channels_last) {
    FGS_REQUIRE(C % 4 == 0, "grid_resize: channels_last needs C % 4 == 0");
    FGS_REQUIRE(((((uintptr_t)src) | ((uintptr_t)dst)) & 15) == 0, "grid_resize: buffers must be 16-byte aligned");
    const int y_strips = fgs::cdiv(Y2, RC_YB);
    const long long n = (long long)X2 * y_strips;
    resize_cl_kernel<<<(unsigned)(n < cap ? n : cap), kBlock, 0, st>>>(src, dst, C, X, Y, Z, X2, Y2, Z2, sx, sy, sz,
                                                                       y_strips, n);
    return fgs::check_launch("grid_resize_cl", st);
  }
  const int y_strips = fgs::cdiv(Y2, RS_YB);
  const long long n = (long long)C * X2 * y_strips;
  const unsigned grid = (unsigned)(n < cap ? n : cap);
  if (Z2 % 4 == 0 && (((uintptr_t)dst) & 15) == 0)
    resize_ncdhw_kernel<4><<<grid, kBlock, 0, st>>>(src, dst, C, X, Y, Z, X2, Y2, Z2, sx, sy, sz, y_strips, n);
  else
    resize_ncdhw_kernel<1><<<grid, kBlock, 0, st>>>(src, dst, C, X, Y, Z, X2, Y2, Z2, sx, sy, sz, y_strips, n);
  return fgs::check_launch("grid_resize", st);
}

extern "C" int fgs_gather_rays(const float* rays_o_tr, const float* rays_d_tr, const float* viewdirs_tr,
                               const float* rgb_tr, int64_t n_total, const int64_t* sel, int n, float* rays_o,
                               float* rays_d, float* viewdirs, float* target, int32_t* bad, fgs_stream_t stream) {
  FGS_REQUIRE(n >= 0 && n_total >= 0, "gather_rays: negative size");
  if (n == 0) return 0;
  FGS_REQUIRE(rays_o_tr && rays_d_tr && viewdirs_tr && rgb_tr && sel && rays_o && rays_d && viewdirs && target,
              "gather_rays: null pointer");
  FGS_REQUIRE(n <= (1 << 28), "gather_rays: batch too large");
  cudaStream_t st = (cudaStream_t)stream;
  gather_rays_kernel<<<fgs::cdiv(4ll * n, kBlock), kBlock, 0, st>>>(rays_o_tr, rays_d_tr, viewdirs_tr, rgb_tr, n_total,
                                                                    sel, n, rays_o, rays_d, viewdirs, target, bad);
  return fgs::check_launch("gather_rays", st);
}

extern "C" int fgs_conv3d_replicate_bwd(const float* g_out, float* g_in, int X, int Y, int Z, const float* weights, int ksize,
                                        fgs_stream_t stream) {
  FGS_REQUIRE(X > 0 && Y > 0 && Z > 0, "conv3d_replicate_bwd: sizes must be positive");
  FGS_REQUIRE(ksize == 1 || ksize == 3 || ksize == 5 || ksize == 7, "conv3d_replicate_bwd: ksize must be 1, 3, 5 or 7");
  FGS_REQUIRE(g_out && g_in && g_out != g_in && weights, "conv3d_replicate_bwd: g_out / g_in must be distinct non-null buffers");
  cudaStream_t st = (cudaStream_t)stream;
  switch (ksize) {
    case 1: return launch_conv_bwd<1>(g_out, g_in, X, Y, Z, weights, st);
    case 3: return launch_conv_bwd<3>(g_out, g_in, X, Y, Z, weights, st);
    case 5: return launch_conv_bwd<5>(g_out, g_in, X, Y, Z, weights, st);
    default: return launch_conv_bwd<7>(g_out, g_in, X, Y, Z, weights, st);
  }
}
