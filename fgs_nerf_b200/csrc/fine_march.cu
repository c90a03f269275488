// Fused ray marching for the fine stage (sm_100a): sampling + MaskCache + SDF / gradient /
// NeuS alpha + segmented transmittance scan, with every data-dependent count kept on the device.
//
// Replaces, per training step, the reference's chain (model/nerf.py:776-833)
//   sample_pts_on_rays (6 launches, 2 scans, sync)  ->  3 masked selects  ->  MaskCache
//   grid_sample -> 3 masked selects -> grid_sample x7 + ~15 elementwise kernels (alpha) ->
//   7 masked selects -> alpha2weight (2 launches + sync) -> 8 masked selects
// by four launches and no host synchronisation:
//   march_kernel   warp per ray: t_min/t_max/N_steps, per-step bbox + MaskCache test, survivors
//                  recorded as a bitmap (1 bit per step) + count
//   scan_kernel    exclusive scan of the per-ray counts (single CTA; N <= 2^20 rays)
//   expand1 / sample1 / weights1  survivors get sdf, 6-tap gradient, alpha (thread per sample); the
//                  transmittance recurrence runs over a warp per ray with shuffles (early stop T<1e-3,
//                  alpha and weight thresholds) and leaves flags + the per-ray level-2 count
//   compact2_kernel warp per ray: level-2 index list
// Arithmetic follows the reference op by op (see fine_common.cuh / ray_math.cuh) so the
// surviving sample sets are bit-identical.
#include "fine_common.cuh"

namespace {

using namespace fine;
constexpr int kBlock = 256;
constexpr int kWarps = kBlock / 32;

// workspace overflow: the per-forward bit AND the sticky word that only the host clears (an asynchronous training
// loop reads the counters every i_print iterations; the per-forward word only ever shows the last one)
__device__ __forceinline__ void flag_overflow(int* counters, int bit) {
  atomicOr(counters + FGS_CNT_OVERFLOW, bit);
  atomicOr(counters + FGS_CNT_STICKY, bit);
}

// march_kernel, warp per ray, kMarchWords bitmap words (32 steps each) of the ray in flight: the steps of a ray are
// independent, but a warp that finishes one word (occupancy word -> 8 mask gathers -> ballot) before it starts the next
// pays two dependent memory latencies per word.  The occupancy loads of the words are issued together, then the gathers,
// then the ballots in step order.  Measured on the headline workload (profiles/r02_variant_sweep.jsonl): 38.1 us with one
// word at a time (64 registers) -> 34.9 us with two words at 40 registers (6 CTAs per SM); four words need > 64
// registers or spill and are no faster (36.9 us), and no register bound alone moves the kernel (35.7-38.3 us).
constexpr int kMarchWords = 2;
__global__ void __launch_bounds__(kBlock, 6) march_kernel(const __grid_constant__ FgsFineParams P) {
  constexpr int W = kMarchWords;
  const int lane = lane_id();
  const GridGeo mg = make_geo(P.mask_xyz_min, P.mask_xyz_max, P.mX, P.mY, P.mZ);
  const float thres = P.mask_thres;
  const uint32_t* __restrict__ occ = (thres > 0.f && P.mask_grid != nullptr) ? P.mask_occ : nullptr;
  long long tot_steps = 0;
  int tot_m0 = 0;
  for (int ray = blockIdx.x * kWarps + (threadIdx.x >> 5); ray < P.n_rays; ray += gridDim.x * kWarps) {
    const float* o = P.rays_o + 3 * (size_t)ray;
    const float* d = P.rays_d + 3 * (size_t)ray;
    const RayIsect is = ray_aabb(o[0], o[1], o[2], d[0], d[1], d[2], P.xyz_min[0], P.xyz_min[1],
                                 P.xyz_min[2], P.xyz_max[0], P.xyz_max[1], P.xyz_max[2], P.near_, P.far_);
    const long long n64 = ray_nsamples(is.t_min, is.t_max, ray_norm(d[0], d[1], d[2]), P.stepdist);
    const int cap = P.cap_words * 32;
    const int n = (int)(n64 < (long long)cap ? n64 : (long long)cap);
    if (lane == 0) {
      P.t_min[ray] = is.t_min;
      P.t_max[ray] = is.t_max;
      P.n_steps[ray] = (int)(n64 < 0x7fffffffLL ? n64 : 0x7fffffffLL);
      if (n64 > cap) flag_overflow(P.counters, 1);
      tot_steps += n64;
    }
    const Ray r = make_ray(o, d, is.t_min);
    int c1 = 0;
    const int words = (n + 31) >> 5;
    for (int w0 = 0; w0 < words; w0 += W) {
      bool in[W], keep[W], gather[W];
      float px[W], py[W], pz[W], iw[W], ih[W], id[W];
      int cell[W];
      // 1. points, bbox test (render_utils_kernel.cu:191-192), MaskCache index, occupancy cell
#pragma unroll
      for (int u = 0; u < W; ++u) {
        const int step = (w0 + u) * 32 + lane;
        in[u] = step < n;
        ray_point(r, P.stepdist, step, px[u], py[u], pz[u]);
        in[u] = in[u] && !((P.xyz_min[0] > px[u]) | (P.xyz_min[1] > py[u]) | (P.xyz_min[2] > pz[u]) |
                           (P.xyz_max[0] < px[u]) | (P.xyz_max[1] < py[u]) | (P.xyz_max[2] < pz[u]));
        keep[u] = in[u] && P.mask_grid == nullptr;           // geometry_searching: no MaskCache
        gather[u] = in[u] && P.mask_grid != nullptr;
        cell[u] = -1;
        iw[u] = ih[u] = id[u] = 0.f;
        if (gather[u]) {
          cont_index(mg, px[u], py[u], pz[u], iw[u], ih[u], id[u]);   // MaskCache.forward, model/nerf.py:1202-1209
          if (occ) {
            const int cx = (int)floorf(iw[u]), cy = (int)floorf(ih[u]), cz = (int)floorf(id[u]);
            if (cx >= 0 && cx < mg.Z - 1 && cy >= 0 && cy < mg.Y - 1 && cz >= 0 && cz < mg.X - 1)
              cell[u] = (int)(((unsigned)cz * (unsigned)(mg.Y - 1) + (unsigned)cy) * (unsigned)(mg.Z - 1) + (unsigned)cx);
          }
        }
      }
      // 2. occupancy words (all 8 corners zero -> the trilinear value is exactly 0 < thres: no gathers)
      uint32_t ow[W];
#pragma unroll
      for (int u = 0; u < W; ++u) ow[u] = (cell[u] >= 0) ? __ldg(occ + ((unsigned)cell[u] >> 5)) : 0xffffffffu;
#pragma unroll
      for (int u = 0; u < W; ++u) gather[u] = gather[u] && ((ow[u] >> (cell[u] & 31)) & 1u);
      // 3. the mask gathers
#pragma unroll
      for (int u = 0; u < W; ++u) {
        if (gather[u]) {
          const Tri t = tri_setup(iw[u], ih[u], id[u]);
          keep[u] = tri_sample1(P.mask_grid, t, mg.X, mg.Y, mg.Z) >= thres;
        }
      }
#pragma unroll
      for (int u = 0; u < W; ++u) {
        if (keep[u] && P.inc_mask != nullptr) {
          // Incremental Voxel Training: MaskGrid.forward = maskcache_lookup (render_utils_kernel.cu:374-392)
          const int i = (int)roundf(__fmaf_rn(px[u], P.inc_scale[0], P.inc_shift[0]));
          const int j = (int)roundf(__fmaf_rn(py[u], P.inc_scale[1], P.inc_shift[1]));
          const int k = (int)roundf(__fmaf_rn(pz[u], P.inc_scale[2], P.inc_shift[2]));
          keep[u] = 0 <= i && i < P.iX && 0 <= j && j < P.iY && 0 <= k && k < P.iZ &&
                    P.inc_mask[((size_t)i * P.iY + j) * P.iZ + k] != 0;
        }
      }
      // 4. bitmap words in step order
#pragma unroll
      for (int u = 0; u < W; ++u) {
        const unsigned m0 = __ballot_sync(0xffffffffu, in[u]);
        const unsigned word = __ballot_sync(0xffffffffu, keep[u]);
        if (lane == 0 && w0 + u < words) {
          P.bits[(size_t)ray * P.cap_words + w0 + u] = word;
          tot_m0 += __popc(m0);
        }
        c1 += __popc(word);
      }
    }
    if (lane == 0) P.cnt1[ray] = c1;
  }
  // one atomic per CTA and counter (per-warp atomics into the same two addresses serialise in L2)
  __shared__ long long s_steps[kWarps];
  __shared__ int s_m0[kWarps];
  if (lane == 0) { s_steps[threadIdx.x >> 5] = tot_steps; s_m0[threadIdx.x >> 5] = tot_m0; }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long ts = 0;
    int tm = 0;
#pragma unroll
    for (int i = 0; i < kWarps; ++i) { ts += s_steps[i]; tm += s_m0[i]; }
    if (tm) atomicAdd(P.counters + FGS_CNT_M0, tm);
    if (ts) atomicAdd((unsigned long long*)(P.counters + FGS_CNT_TOTAL_LO), (unsigned long long)ts);
  }
}

// Exclusive scan of cnt[0..n) -> off[0..n]; total -> counters[slot]; overflow bit if > cap.
__global__ void __launch_bounds__(1024) scan_kernel(const int* __restrict__ cnt, int* __restrict__ off,
                                                    int n, int* __restrict__ counters, int slot,
                                                    int cap, int overflow_bit) {
  __shared__ int s_warp[32];
  const int per = (n + 1023) / 1024;
  const int beg = threadIdx.x * per;
  const int end = min(beg + per, n);
  // per <= 8 (n <= 8192 rays, the usual batch): the thread's run lives in registers, read once with all loads in
  // flight (the dependent `sum += cnt[i]` loop paid one memory latency per element, twice)
  constexpr int kReg = 8;
  int r[kReg];
  const bool in_regs = per <= kReg;
  int sum = 0;
  if (in_regs) {
#pragma unroll
    for (int i = 0; i < kReg; ++i) r[i] = (i < per && beg + i < n) ? cnt[beg + i] : 0;
#pragma unroll
    for (int i = 0; i < kReg; ++i) sum += r[i];
  } else {
    for (int i = beg; i < end; ++i) sum += cnt[i];
  }
  const int incl = warp_incl_scan(sum);
  const int w = threadIdx.x >> 5;
  if (lane_id() == 31) s_warp[w] = incl;
  __syncthreads();
  if (w == 0) {
    const int v = s_warp[lane_id()];
    const int sc = warp_incl_scan(v);
    s_warp[lane_id()] = sc - v;
    if (lane_id() == 31) {
      off[n] = sc;
      counters[slot] = sc;
      if (sc > cap) flag_overflow(counters, overflow_bit);
    }
  }
  __syncthreads();
  int run = incl - sum + s_warp[w];
  if (in_regs) {
#pragma unroll
    for (int i = 0; i < kReg; ++i) {
      if (i < per && beg + i < n) off[beg + i] = run;
      run += r[i];
    }
  } else {
    for (int i = beg; i < end; ++i) {
      off[i] = run;
      run += cnt[i];
    }
  }
}


// Exclusive scan of the per-ray counts FUSED into the kernels that consume the offsets (expand1, compact2): a CTA owns
// a contiguous range of rays, sums the counts of all rays in front of it (<= 8192 ints: 32 coalesced loads per thread),
// scans its own (<= kRaysPerCta) counts with one warp and publishes off[ray]; the last CTA also publishes the total, the
// device-side count and the overflow bit.  This replaces a single-CTA scan launch (12 us each, twice per forward).
constexpr int kRaysPerCta = 16;        // 512 CTAs at 8192 rays, two rays per warp (64 rays per CTA left 128 CTAs of 8 warps
                                       // walking 8 rays each, one after the other: expand1 18 us, compact2 16 us)
constexpr int kScanPerLane = (kRaysPerCta + 31) / 32;
__device__ __forceinline__ void cta_scan_counts(const int* __restrict__ cnt, int* __restrict__ off, int n, int r0, int r1,
                                                int* __restrict__ counters, int slot, int cap, int overflow_bit,
                                                int* s_off /* [kRaysPerCta + 1] */, int* s_red /* [kWarps] */,
                                                bool prescanned) {
  const int tid = threadIdx.x, lane = lane_id(), w = tid >> 5;
  if (prescanned) {                                    // very large batches: scan_kernel has already written off[0..n]
    if (tid <= kRaysPerCta && r0 + tid <= n) s_off[tid] = off[r0 + tid];
    __syncthreads();
    return;
  }
  int part = 0;
  for (int i = tid; i < r0; i += kBlock) part += cnt[i];
  part = warp_sum(part);
  if (lane == 0) s_red[w] = part;
  __syncthreads();
  if (w == 0) {
    int base = lane < kWarps ? s_red[lane] : 0;
    base = warp_sum(base);
    // the CTA's own rays: kScanPerLane consecutive rays per lane
    int c[kScanPerLane], sum = 0;
#pragma unroll
    for (int j = 0; j < kScanPerLane; ++j) {
      const int local = lane * kScanPerLane + j, i = r0 + local;
      c[j] = (local < kRaysPerCta && i < r1) ? cnt[i] : 0;
      sum += c[j];
    }
    const int incl = warp_incl_scan(sum);
    int run = base + incl - sum;
#pragma unroll
    for (int j = 0; j < kScanPerLane; ++j) {
      const int local = lane * kScanPerLane + j, i = r0 + local;
      if (local < kRaysPerCta) s_off[local] = run;
      if (local < kRaysPerCta && i < r1) off[i] = run;
      run += c[j];
    }
    if (lane == 31) {
      s_off[kRaysPerCta] = base + incl;
      if (r1 == n) {                                   // last CTA: the total
        const int total = base + incl;
        off[n] = total;
        counters[slot] = total;
        if (total > cap) flag_overflow(counters, overflow_bit);
      }
    }
  }
  __syncthreads();
}

// trilinear SDF value + 6-tap finite-difference gradient (nerf.grid_sampler, model/nerf.py:
// 639-672 with sample_sdfs(displace=[1.0]) :597-637); gradient returned in x,y,z order (:664).
__device__ __forceinline__ void sdf_and_grad(const FgsFineParams& P, const GridGeo& g, float px, float py,
                                             float pz, float& sdf, float& gx, float& gy, float& gz) {
  float iw, ih, id;
  cont_index(g, px, py, pz, iw, ih, id);
  const Tri t0 = tri_setup(iw, ih, id);
  sdf = tri_sample1(P.sdf, t0, g.X, g.Y, g.Z);
  float gr[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const TapPair tp = make_taps(g, iw, ih, id, a, 1.0f);
    const float flo = tri_sample1(P.sdf, tp.lo, g.X, g.Y, g.Z);
    const float fhi = tri_sample1(P.sdf, tp.hi, g.X, g.Y, g.Z);
    gr[a] = tap_grad(flo, fhi, tp.diff, P.voxel_size);
  }
  gx = gr[2]; gy = gr[1]; gz = gr[0];
}

// Level 1 runs as three kernels so that the expensive part is balanced over SAMPLES, not rays (rays keep
// between 0 and ~200 samples after the MaskCache; with a warp walking one ray's step bitmap the longest rays
// set the kernel time while 90 % of the warp slots sat empty: 85 us, sm__warps_active 11 %):
//   expand1_kernel   warp per ray: step bitmap -> (ray_id1, step_id1) at off1[ray] + rank
//   sample1_kernel   thread per level-1 sample: point, SDF, 6-tap gradient, NeuS alpha (56 gathers)
//   weights1_kernel  warp per ray: transmittance recurrence over the ray's alphas (early stop T < 1e-3,
//                    alpha / weight thresholds), flags, level-2 count
// The arithmetic per sample and the order of the recurrence are unchanged.
__global__ void __launch_bounds__(kBlock) expand1_kernel(const __grid_constant__ FgsFineParams P, bool prescanned) {
  __shared__ int s_off[kRaysPerCta + 1], s_red[kWarps];
  const int lane = lane_id();
  const int r0 = blockIdx.x * kRaysPerCta, r1 = min(r0 + kRaysPerCta, P.n_rays);
  cta_scan_counts(P.cnt1, P.off1, P.n_rays, r0, r1, P.counters, FGS_CNT_M1, P.cap_m1, 2, s_off, s_red, prescanned);
  for (int ray = r0 + (threadIdx.x >> 5); ray < r1; ray += kWarps) {
    const int base = s_off[ray - r0];
    const int c1 = P.cnt1[ray];
    if (c1 == 0) continue;
    if (base + c1 > P.cap_m1) {           // over capacity (flagged by the scan): mark the slots that exist
      for (int i = base + lane; i < min(base + c1, P.cap_m1); i += 32) P.ray_id1[i] = -1;
      continue;
    }
    int n = P.n_steps[ray];
    n = min(n, P.cap_words * 32);
    const int words = (n + 31) >> 5;
    int run = 0;
    for (int w = 0; w < words; ++w) {
      const unsigned word = P.bits[(size_t)ray * P.cap_words + w];
      if (word == 0) continue;
      if ((word >> lane) & 1u) {
        const int idx = base + run + __popc(word & lanemask_lt());
        P.ray_id1[idx] = ray;
        P.step_id1[idx] = w * 32 + lane;
      }
      run += __popc(word);
    }
  }
}

__global__ void __launch_bounds__(kBlock) sample1_kernel(const __grid_constant__ FgsFineParams P) {
  const GridGeo g = make_geo(P.xyz_min, P.xyz_max, P.X, P.Y, P.Z);
  const int M1 = min(P.counters[FGS_CNT_M1], P.cap_m1);
  for (int i = blockIdx.x * kBlock + threadIdx.x; i < M1; i += gridDim.x * kBlock) {
    const int ray = P.ray_id1[i];
    if (ray < 0) continue;
    const int step = P.step_id1[i];
    const float* o = P.rays_o + 3 * (size_t)ray;
    const float* d = P.rays_d + 3 * (size_t)ray;
    const float vx = P.viewdirs[3 * (size_t)ray], vy = P.viewdirs[3 * (size_t)ray + 1],
                vz = P.viewdirs[3 * (size_t)ray + 2];
    const Ray r = make_ray(o, d, P.t_min[ray]);
    float px, py, pz, sdf, gx, gy, gz;
    ray_point(r, P.stepdist, step, px, py, pz);
    sdf_and_grad(P, g, px, py, pz, sdf, gx, gy, gz);
    P.sdf1[i] = sdf;
    P.grad1[i] = gx;
    P.grad1[(size_t)P.cap_m1 + i] = gy;
    P.grad1[2 * (size_t)P.cap_m1 + i] = gz;
    P.alpha1[i] = neus_alpha(sdf, gx, gy, gz, vx, vy, vz, P.dist, P.inv_s);
  }
}

__global__ void __launch_bounds__(kBlock) weights1_kernel(const __grid_constant__ FgsFineParams P) {
  const int lane = lane_id();
  const float thres = P.fast_color_thres;
  int tot_a = 0;
  for (int ray = blockIdx.x * kWarps + (threadIdx.x >> 5); ray < P.n_rays; ray += gridDim.x * kWarps) {
    const int base = P.off1[ray];
    const int c1 = P.cnt1[ray];
    if (c1 == 0 || base + c1 > P.cap_m1) {  // empty ray, or over capacity (flagged by the scan)
      if (lane == 0) { P.cnt2[ray] = 0; P.alphainv_last[ray] = 1.f; }
      continue;
    }
    float Tc = 1.f;
    bool stopped = false;
    int c2 = 0;
    for (int b = 0; b < c1; b += 32) {
      const bool mine = b + lane < c1;
      const int idx = base + b + lane;
      const float alpha = mine ? P.alpha1[idx] : 0.f;
      const bool kept = mine && (alpha > thres);  // model/nerf.py:803
      const float a_eff = kept ? alpha : 0.f;
      unsigned m = __ballot_sync(0xffffffffu, kept);
      tot_a += __popc(m);
      float myT = 1.f, myW = 0.f;
      bool visited = false;
      // render_utils_kernel.cu:577-605 over the level-a samples of this chunk, in order
      while (m && !stopped) {
        const int j = __ffs(m) - 1;
        m &= m - 1;
        const float aj = __shfl_sync(0xffffffffu, a_eff, j);
        if (lane == j) {
          myT = Tc;
          myW = __fmul_rn(Tc, aj);
          visited = true;
        }
        Tc = (float)__dmul_rn(__dsub_rn(1., (double)aj), (double)Tc);
        if ((double)Tc < 1e-3) stopped = true;
      }
      const bool l2 = visited && (myW > thres);  // model/nerf.py:825
      if (mine) {
        P.w1[idx] = myW;
        P.T1[idx] = myT;
        P.flag1[idx] = (kept ? FGS_F_ALPHA : 0) | (visited ? FGS_F_VISIT : 0) | (l2 ? FGS_F_L2 : 0);
      }
      c2 += __popc(__ballot_sync(0xffffffffu, l2));
    }
    if (lane == 0) {
      P.cnt2[ray] = c2;
      P.alphainv_last[ray] = Tc;
    }
  }
  __shared__ int s_a[kWarps];
  if (lane == 0) s_a[threadIdx.x >> 5] = tot_a;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
#pragma unroll
    for (int i = 0; i < kWarps; ++i) t += s_a[i];
    if (t) atomicAdd(P.counters + FGS_CNT_MA, t);
  }
}

__global__ void __launch_bounds__(kBlock) compact2_kernel(const __grid_constant__ FgsFineParams P, bool prescanned) {
  __shared__ int s_off[kRaysPerCta + 1], s_red[kWarps];
  const int lane = lane_id();
  const int r0 = blockIdx.x * kRaysPerCta, r1 = min(r0 + kRaysPerCta, P.n_rays);
  cta_scan_counts(P.cnt2, P.off2, P.n_rays, r0, r1, P.counters, FGS_CNT_M2, P.cap_m2, 4, s_off, s_red, prescanned);
  for (int ray = r0 + (threadIdx.x >> 5); ray < r1; ray += kWarps) {
    const int c2 = P.cnt2[ray];
    if (c2 == 0) continue;
    const int base1 = P.off1[ray], c1 = P.cnt1[ray];
    int dst = s_off[ray - r0];
    // a ray that does not fit (flagged by the scan: the step's results are void) still fills the slots that exist, so
    // that every row below min(M2, cap_m2) names a real sample and nothing downstream indexes with stale memory
    if (dst >= P.cap_m2) continue;
    for (int b = 0; b < c1; b += 32) {
      const int i = b + lane;
      const bool l2 = (i < c1) && (P.flag1[base1 + i] & FGS_F_L2);
      const unsigned m = __ballot_sync(0xffffffffu, l2);
      if (l2) {
        const int s = dst + __popc(m & lanemask_lt());
        if (s < P.cap_m2) {
          P.idx1_of2[s] = base1 + i;
          P.ray_id2[s] = ray;
        }
      }
      dst += __popc(m);
    }
  }
}

// one bit per cell (z0, y0, x0 = the floor corner in D, H, W order) of the MaskCache grid: any of its 8 corners non-zero
__global__ void __launch_bounds__(kBlock)
mask_occupancy_kernel(const float* __restrict__ grid, int D, int H, int W, uint32_t* __restrict__ occ) {
  const long long cells = (long long)(D - 1) * (H - 1) * (W - 1);
  for (long long c = blockIdx.x * (long long)kBlock + threadIdx.x; c < cells; c += (long long)gridDim.x * kBlock) {
    long long r = c;
    const int x = (int)(r % (W - 1)); r /= (W - 1);
    const int y = (int)(r % (H - 1)); r /= (H - 1);
    const int z = (int)r;
    bool any = false;
#pragma unroll
    for (int k = 0; k < 8; ++k)
      any |= grid[((size_t)(z + (k >> 2)) * H + (y + ((k >> 1) & 1))) * W + (x + (k & 1))] != 0.f;
    if (any) atomicOr(occ + (c >> 5), 1u << (c & 31));
  }
}

}  // namespace

extern "C" int64_t fgs_mask_occupancy_words(int mX, int mY, int mZ) {
  if (mX < 2 || mY < 2 || mZ < 2) return 1;
  return ((int64_t)(mX - 1) * (mY - 1) * (mZ - 1) + 31) / 32;
}

extern "C" int fgs_mask_occupancy(const float* mask_grid, int mX, int mY, int mZ, uint32_t* occ, fgs_stream_t stream) {
  FGS_REQUIRE(mask_grid && occ && mX >= 2 && mY >= 2 && mZ >= 2, "mask_occupancy: grid must be at least 2^3");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t words = fgs_mask_occupancy_words(mX, mY, mZ);
  int rc = fgs::check_cuda(cudaMemsetAsync(occ, 0, sizeof(uint32_t) * (size_t)words, st), "mask_occupancy memset");
  if (rc) return rc;
  const long long cells = (long long)(mX - 1) * (mY - 1) * (mZ - 1);
  mask_occupancy_kernel<<<fgs::stream_grid(cells, kBlock, 8), kBlock, 0, st>>>(mask_grid, mX, mY, mZ, occ);
  return fgs::check_launch("mask_occupancy", st);
}

namespace fine {

// sampling + MaskCache + level-1 id arrays only (shared with the coarse stage, whose per-sample kernels differ)
int launch_march_expand(const FgsFineParams& P, cudaStream_t st) {
  int rc = fgs::check_cuda(cudaMemsetAsync(P.counters, 0, sizeof(int32_t) * FGS_CNT_STICKY, st),
                           "fine counters memset");
  if (rc) return rc;
  fgs::prof_mark("memset", st);
  const int grid = fgs::stream_grid(P.n_rays, kWarps, 8);
  march_kernel<<<grid, kBlock, 0, st>>>(P);
  if ((rc = fgs::check_launch("fine march", st))) return rc;
  const int grid_r = fgs::cdiv(P.n_rays, kRaysPerCta);
  const bool prescan = P.n_rays > 32768;
  if (prescan) {
    scan_kernel<<<1, 1024, 0, st>>>(P.cnt1, P.off1, P.n_rays, P.counters, FGS_CNT_M1, P.cap_m1, 2);
    if ((rc = fgs::check_launch("fine scan1", st))) return rc;
  }
  expand1_kernel<<<grid_r, kBlock, 0, st>>>(P, prescan);
  return fgs::check_launch("fine expand1", st);
}

// level-2 index list from the flags the weights kernel left (shared with the coarse stage)
int launch_compact2(const FgsFineParams& P, cudaStream_t st) {
  int rc = 0;
  const int grid_r = fgs::cdiv(P.n_rays, kRaysPerCta);
  const bool prescan = P.n_rays > 32768;
  if (prescan) {
    scan_kernel<<<1, 1024, 0, st>>>(P.cnt2, P.off2, P.n_rays, P.counters, FGS_CNT_M2, P.cap_m2, 4);
    if ((rc = fgs::check_launch("fine scan2", st))) return rc;
  }
  compact2_kernel<<<grid_r, kBlock, 0, st>>>(P, prescan);
  return fgs::check_launch("fine compact2", st);
}

int launch_march(const FgsFineParams& P, cudaStream_t st) {
  int rc = fgs::check_cuda(cudaMemsetAsync(P.counters, 0, sizeof(int32_t) * FGS_CNT_STICKY, st),
                           "fine counters memset");
  if (rc) return rc;
  fgs::prof_mark("memset", st);
  const int grid = fgs::stream_grid(P.n_rays, kWarps, 8);
  march_kernel<<<grid, kBlock, 0, st>>>(P);
  if ((rc = fgs::check_launch("fine march", st))) return rc;
  const int grid_r = fgs::cdiv(P.n_rays, kRaysPerCta);          // scan fused: a CTA owns kRaysPerCta consecutive rays
  // (every CTA re-reads the counts in front of it: O(n^2 / 64) loads, fine up to the usual 8192-ray batches; beyond
  // 32 K rays the single-CTA scan runs first)
  const bool prescan = P.n_rays > 32768;
  if (prescan) {
    scan_kernel<<<1, 1024, 0, st>>>(P.cnt1, P.off1, P.n_rays, P.counters, FGS_CNT_M1, P.cap_m1, 2);
    if ((rc = fgs::check_launch("fine scan1", st))) return rc;
  }
  expand1_kernel<<<grid_r, kBlock, 0, st>>>(P, prescan);
  if ((rc = fgs::check_launch("fine expand1", st))) return rc;
  // the sample count lives on the device: size the grid for the capacity, capped at a few waves (grid-stride loop)
  const int grid_s = fgs::stream_grid(P.cap_m1, kBlock, 8);
  sample1_kernel<<<grid_s, kBlock, 0, st>>>(P);
  if ((rc = fgs::check_launch("fine level1", st))) return rc;
  weights1_kernel<<<grid, kBlock, 0, st>>>(P);
  if ((rc = fgs::check_launch("fine weights1", st))) return rc;
  if (prescan) {
    scan_kernel<<<1, 1024, 0, st>>>(P.cnt2, P.off2, P.n_rays, P.counters, FGS_CNT_M2, P.cap_m2, 4);
    if ((rc = fgs::check_launch("fine scan2", st))) return rc;
  }
  compact2_kernel<<<grid_r, kBlock, 0, st>>>(P, prescan);
  return fgs::check_launch("fine compact2", st);
}

}  // namespace fine
