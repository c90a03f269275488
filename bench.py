#!/usr/bin/env python
"""bench.py -- train rays/s of the fused fine-stage step (fwd + bwd + TV + masked Adam).

    python bench.py --gpus N --steps K --warmup W            (our arm; torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K --warmup W   (reference's CPU path)
    python bench.py --grid 256 --camera dtu ...               (BASELINE configs 3 / 4)
    python bench.py --mode render --grid 384 ...              (BASELINE config 5: full-image inference)
    python bench.py --stage coarse --grid 116 --global-step 3000   (the coarse-stage iteration)

Workload (BASELINE.json configs[1], the default): shiny-blender fine stage, 160^3 SDF + 12-channel feature grid,
8192 rays per GPU per step, <= 256 post-mask samples per ray, synthetic scene of SURVEY.md 8d (random-init grids +
MLPs, blender pinhole rays, random targets), global_step 5000 onwards (TV regularisers on every 3rd step, as
config/shiny_blender.py:180-219 schedules them).

The trajectory is FROZEN by default (--freeze 1): all learning rates are 0, so every step launches the same kernels
on the same bytes (Adam still reads and writes p, m, v and the gradients) while the sample counts stay what the scene
gives at initialisation -- with real rates Adam fits the random targets and M2 per ray falls by a third within one
run, which made the regions of one run incomparable.

One JSON line on stdout (rank 0):
  value      rays/s, inputs resident in HBM, K steps timed with CUDA events, max over ranks
  e2e        same metric through the public host API: rays + targets start in pinned HOST memory,
             H2D copies and a D2H read of the step's loss + sample counters inside the timed region
  roofline   the dominant kernel of the step, timed per launch with CUDA events (fgs_prof_*)
  kernels    every launch of the step with its byte / FLOP model and fraction of the measured peak
  cpu_baseline  the oracle (torch CPU restatement of the reference step) on a bounded ray sample
  gpu_oracle    the oracle on this GPU: torch CUDA ops + the reference's own compiled kernels (oracle/_ref)
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "train_rays_per_sec"
UNIT = "rays/s"
MLP_MACS = 106 * 256 + 3 * 256 * 256 + 307 * 256 + 2 * 256 * 256 + 3 * 256   # per sample, forward


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tensor=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, tensor=1400.0, src="fallback")


def ncu_traffic():
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the kernels captured with
    `ncu --set full` for this round, written by tools/ncu_digest.py into profiles/ next to the captures."""
    p = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
    if os.path.exists(p):
        return json.load(open(p))
    return {}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons while the timed region runs (B200_PROFILING.md):
    one nvidia-smi process in loop mode (-lms 20), started before and killed after the region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            time.sleep(0.3)          # let it take its first samples before the region starts
        except Exception:
            self.proc = None

    def start(self):
        self.t0 = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            out = self.proc.communicate(timeout=5)[0]
        except Exception:
            self.proc.kill()
            out = ""
        rows = [[x.strip() for x in l.split(",")] for l in out.strip().splitlines()]
        rows = [r for r in rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(int(float(r[0])) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(float(rows[0][1])), "reasons": reasons,
                "samples": len(rows)}


class E2EPipe:
    """The end-to-end leg as a host-fed training loop runs it: the inputs of step i + 1 are uploaded (ONE copy of a
    pinned [4, N, 3] block: rays_o, rays_d, viewdirs, target) on a copy stream while step i computes, and the host reads
    step i's loss + sample counters (a D2H copy enqueued right behind the step) once step i + 1 has been enqueued.  Every
    step's inputs cross the bus and every step's result is read by the host, all inside the timed region; only the
    waiting is overlapped.  `serial=True` is the unpipelined loop (copy, step, read, synchronise) for comparison."""

    def __init__(self, torch, host, dev, n_loss, train_step, serial=False):
        self.torch, self.host, self.serial, self.train_step = torch, host, serial, train_step
        self.stg = [torch.empty_like(host[0], device=dev) for _ in range(2)]
        self.loss_h = [torch.empty(n_loss, dtype=torch.float64).pin_memory() for _ in range(2)]
        self.cnt_h = [torch.empty(8, dtype=torch.int32).pin_memory() for _ in range(2)]
        self.copy_s = torch.cuda.Stream(device=dev)
        self.ready = [torch.cuda.Event() for _ in range(2)]
        self.done = [torch.cuda.Event() for _ in range(2)]
        self.primed = None
        self.last = 0

    def _prefetch(self, i):
        b = i % 2
        with self.torch.cuda.stream(self.copy_s):
            self.copy_s.wait_event(self.done[b])          # the step that last read this buffer has finished
            self.stg[b].copy_(self.host[i % len(self.host)], non_blocking=True)
            self.ready[b].record(self.copy_s)
        self.primed = i

    def step(self, i, eng):
        torch = self.torch
        b = i % 2
        main = torch.cuda.current_stream()
        if self.serial:
            self.stg[b].copy_(self.host[i % len(self.host)], non_blocking=True)
        else:
            if self.primed != i:
                self._prefetch(i)
            main.wait_event(self.ready[b])
        s = self.stg[b]
        self.train_step(s[0], s[1], s[2], s[3], i)
        self.loss_h[b].copy_(eng.ws["loss"], non_blocking=True)
        self.cnt_h[b].copy_(eng.ws["counters"], non_blocking=True)
        self.done[b].record(main)
        self.last = b
        if self.serial:
            main.synchronize()                            # the caller reads the step's loss
        else:
            self._prefetch(i + 1)
            self.done[1 - b].synchronize()                # the previous step's loss and counters are on the host

    def loss(self):
        self.torch.cuda.synchronize()
        return self.loss_h[self.last]



def camera(args):
    """blender: 800x800 pinhole (config/shiny_blender.py); dtu: inverse_y, 800x600 (config/dtu.py)."""
    if args.camera == "dtu":
        return dict(H=600, W=800, inverse_y=True)
    return dict(H=800, W=800, inverse_y=False)


def workload(args, rays=None):
    rays = args.rays if rays is None else rays
    if args.mode == "render":
        return {"workload": f"full-image render (inference, render_depth + render_grad), {args.grid}^3 SDF + 12-ch k0 "
                            f"grid, {args.image_h}x{args.image_w} rays per image in {args.rays}-ray chunks dealt "
                            f"round-robin to the GPUs, stepsize 0.5, global_step {args.global_step}",
                "grid": args.grid, "rays_per_image": args.image_h * args.image_w, "chunk": args.rays,
                "camera": args.camera, "global_step": args.global_step,
                "l2_policy": "k0 grid (48 B/voxel) larger than the 126 MB L2; every image is a different view"}
    return {"workload": f"shiny-blender fine-stage training step (fwd+bwd+TV every 3rd+masked Adam), "
                        f"{args.grid}^3 SDF + 12-ch k0 grid, {rays} rays/GPU/step, stepsize 0.5, "
                        f"{args.camera} camera, global_step {args.global_step}+",
            "grid": args.grid, "rays_per_gpu": rays, "camera": args.camera, "global_step": args.global_step,
            "frozen_trajectory": bool(args.freeze),
            "l2_policy": f"working set (k0 + grad + Adam state = 4 x {48 * args.grid ** 3 / 1e6:.0f} MB) larger than the "
                         f"126 MB L2"}


# ---------------------------------------------------------------------------------------------
def _oracle_setup(args, device, n, freeze):
    import torch
    from fgs_nerf_b200 import scene
    from oracle import native_ref, nerf_ref as R
    sc = scene.make_scene(G=args.grid)
    m = R.FineModel(sc, device=device) if device is not None else R.FineModel(sc)
    m.set_nonempty_mask()
    if device is None:
        native_ref.build()
        nat = R.CpuNative()
    else:
        nat = R.RefCudaNative()
    opt = R.RefOptimizer(m, scene.FINE_TRAIN, nat)
    if freeze:
        for g in opt.groups:
            g["lr"] = 0.0
    return sc, m, nat, opt, R


def run_reference(args):
    """Reference arm: the reference's own CPU path for this step = oracle/nerf_ref.py (the torch
    restatement pinned against model/nerf.py) + oracle/fgs_oracle.c for the three native ops, on
    all host threads, a bounded ray sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from fgs_nerf_b200 import scene
    # torchrun exports OMP_NUM_THREADS=1 to its workers: this arm is ONE process, give it the host's cores back
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n = args.ref_rays
    sc, m, nat, opt, R = _oracle_setup(args, None, n, args.freeze)
    cam = camera(args)
    batches = [scene.make_rays(N=n, offset=b * n, **cam) for b in range(2)]
    times = []
    for i in range(args.warmup + args.steps):
        ro, rd, vd, tgt = batches[i % len(batches)]
        t0 = time.perf_counter()
        R.train_step(m, nat, opt, ro, rd, vd, tgt, args.global_step + i, scene.FINE_TRAIN, near=2.0, stepsize=0.5, bg=1.0)
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    v = n * len(times) / total
    sample = (f"{n} rays/step (a bounded sample of the {args.rays}-ray step) of the same {args.grid}^3 workload, "
              f"{len(times)} steps, {args.warmup} warm-up")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": workload(args, rays=n),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_baseline(args):
    """Oracle step on a bounded sample, timed on the host cores (rank 0, N = 1 only)."""
    import torch
    from fgs_nerf_b200 import scene
    n = 2 * args.ref_rays
    sc, m, nat, opt, R = _oracle_setup(args, None, n, args.freeze)
    ro, rd, vd, tgt = scene.make_rays(N=n, **camera(args))
    times = []
    for i in range(3):
        t0 = time.perf_counter()
        R.train_step(m, nat, opt, ro, rd, vd, tgt, args.global_step + i, scene.FINE_TRAIN, near=2.0, stepsize=0.5, bg=1.0)
        times.append(time.perf_counter() - t0)
    t = min(times[1:])
    return {"value": n / t, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{n} rays of the same workload, best of 2 steps after 1 warm-up ({t:.2f} s/step)"}


def gpu_oracle(args, dev):
    """SURVEY 8d-i / BASELINE.md B1: the reference's arithmetic on THIS GPU -- oracle/nerf_ref.py with torch CUDA ops
    (F.grid_sample, F.linear, index_add_) and the reference's own compiled kernels (oracle/_ref: sample_pts_on_rays,
    alpha2weight, total_variation, adam) -- same scene, same rays, full step incl. optimiser; CUDA events, 5 warm-up
    + 20 timed steps, median.  A reported baseline (kind 'reference kernels + torch ops'), not a product path."""
    import torch
    from oracle import ref_ext
    if not ref_ext.available():
        return {"unavailable": "oracle/_ref not built"}
    from fgs_nerf_b200 import scene
    n = args.rays
    sc, m, nat, opt, R = _oracle_setup(args, dev, n, args.freeze)
    ro, rd, vd, tgt = [t.to(dev) for t in scene.make_rays(N=n, **camera(args))]
    ms = []
    for i in range(25):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        R.train_step(m, nat, opt, ro, rd, vd, tgt, args.global_step + i, scene.FINE_TRAIN, near=2.0, stepsize=0.5, bg=1.0)
        e1.record()
        torch.cuda.synchronize()
        if i >= 5:
            ms.append(e0.elapsed_time(e1))
    ms.sort()
    med = ms[len(ms) // 2]
    del m, opt
    torch.cuda.empty_cache()
    return {"value": n / (med * 1e-3), "unit": UNIT, "ms_per_step": med, "steps": len(ms), "warmup": 5,
            "kind": "oracle on CUDA: torch ops + the reference's compiled render_utils / adam_upd / total_variation kernels",
            "rays_per_step": n, "timing": "CUDA events per step incl. the reference's host syncs, median"}


# ---------------------------------------------------------------------------------------------
def _dilate(occ, lo, hi, axis):
    """Voxels read by trilinear taps displaced by up to -lo .. +hi cells along `axis` (and the 2-cell trilinear
    footprint on the other axes) from the base cells marked in occ [X,Y,Z] bool."""
    import torch.nn.functional as F
    pads = []
    for ax in (2, 1, 0):                          # F.pad takes the last dimension first
        pads += [hi, lo] if ax == axis else [1, 0]
    k = [lo + hi + 1 if ax == axis else 2 for ax in (0, 1, 2)]
    x = F.pad(occ[None, None].float(), pads)
    return F.max_pool3d(x, kernel_size=k, stride=1)[0, 0] > 0


def _cells(p, mn, mx, size):
    import torch
    idx = (p - mn) / (mx - mn) * (size - 1)
    return idx.floor().long().clamp_(min=0).minimum((size - 2).long())


def unique_voxel_counts(eng, ro, rd):
    """Unique grid voxels the gather / scatter kernels touch in one step (their compulsory grid traffic):
    recomputed from the engine's workspace after a forward with torch ops (bench-only, untimed)."""
    import torch
    from fgs_nerf_b200.engine import F_VISIT
    dev = eng.device
    ws = eng.ws
    cnt = eng.counters()
    M1, M2 = cnt["M1"], cnt["M2"]
    size = torch.tensor([eng.X, eng.Y, eng.Z], device=dev, dtype=torch.float32)
    mn, mx = eng.xyz_min.to(dev), eng.xyz_max.to(dev)
    stepdist = float(eng.P.stepdist)

    def points(ray, step):
        d = rd[ray]
        start = ro[ray] + d * ws["t_min"][ray][:, None]
        return start + d / d.norm(dim=1, keepdim=True) * (stepdist * step.float())[:, None]

    def occupancy(p, shape, mn_, mx_):
        sz = torch.tensor(shape, device=dev, dtype=torch.float32)
        c = _cells(p, mn_, mx_, sz)
        occ = torch.zeros(shape, dtype=torch.bool, device=dev)
        occ[c[:, 0], c[:, 1], c[:, 2]] = True
        return occ

    ray1, step1 = ws["ray_id1"][:M1].long(), ws["step_id1"][:M1]
    p1 = points(ray1, step1)
    shape = (eng.X, eng.Y, eng.Z)
    occ1 = occupancy(p1, shape, mn, mx)
    taps1 = _dilate(occ1, 1, 2, 0) | _dilate(occ1, 1, 2, 1) | _dilate(occ1, 1, 2, 2)
    vis = (ws["flag1"][:M1] & F_VISIT) != 0
    occ1v = occupancy(p1[vis], shape, mn, mx)
    taps1v = _dilate(occ1v, 1, 2, 0) | _dilate(occ1v, 1, 2, 1) | _dilate(occ1v, 1, 2, 2)
    idx2 = ws["idx1_of2"][:M2].long()
    occ2 = occupancy(p1[idx2], shape, mn, mx)
    taps2 = _dilate(occ2, 2, 3, 0) | _dilate(occ2, 2, 3, 1) | _dilate(occ2, 2, 3, 2)
    k0v = _dilate(occ2, 0, 1, 0)                     # the 8 trilinear corners
    # mask grid cells read by the march: every in-bbox step of every ray
    n_steps = ws["n_steps"].long().clamp(max=eng.cap_words * 32)
    ray0 = torch.repeat_interleave(torch.arange(eng.n_rays, device=dev), n_steps)
    first = torch.cumsum(n_steps, 0) - n_steps
    step0 = torch.arange(ray0.numel(), device=dev) - first[ray0]
    p0 = points(ray0, step0)
    inb = ((p0 >= mn) & (p0 <= mx)).all(dim=1)
    mshape = tuple(eng.mask_grid.shape[2:])
    occm = occupancy(p0[inb], mshape, eng.mask_xyz_min.to(dev), eng.mask_xyz_max.to(dev))
    maskv = _dilate(occm, 0, 1, 0)
    return dict(sdf_level1=int(taps1.sum()), sdf_level1_visited=int(taps1v.sum()), sdf_level2=int(taps2.sum()),
                k0_level2=int(k0v.sum()), mask_march=int(maskv.sum()), visited=int(vis.sum()))


def byte_models(cnt, uv, N, V, tv_frac, bricks):
    """Algorithmic (compulsory) work per step of every kernel: unique voxels touched x bytes + the per-sample /
    per-ray arrays the kernel has to read and write once.  name -> (bound, work, unit, model)."""
    M1, M2, total_len = cnt["M1"], cnt["M2"], cnt["total_len"]
    vis = uv["visited"]
    mac = 2.0 * MLP_MACS * M2
    B = {}
    B["mlp_chain_fwd"] = ("tensor", mac / 1e12, "TFLOP/s", "2 x 434176 MAC x M2 (fp32 semantics; 3xTF32 issues 3x)")
    B["mlp_chain_bwd"] = ("tensor", mac / 1e12, "TFLOP/s", "2 x 434176 MAC x M2 (backward-data)")
    B["mlp_dw"] = ("tensor", mac / 1e12, "TFLOP/s", "2 x 434176 MAC x M2 (weight gradients)")
    g = lambda b: b / 1e9
    B["fine march"] = ("hbm", g(24 * N + 16 * N + total_len / 8 + 4 * uv["mask_march"]), "GB/s",
                       "rays 24N + per-ray out 16N + step bitmap total_len/8 + unique mask voxels x 4")
    B["fine expand1"] = ("hbm", g(total_len / 8 + 8 * M1), "GB/s", "bitmap total_len/8 + ids 8 M1")
    B["fine level1"] = ("hbm", g(8 * M1 + 36 * N + 4 * uv["sdf_level1"] + 20 * M1), "GB/s",
                        "ids 8 M1 + rays 36N + unique sdf voxels x 4 + (sdf, grad3, alpha) 20 M1")
    B["fine weights1"] = ("hbm", g(4 * M1 + 9 * M1 + 8 * N), "GB/s", "alpha 4 M1 + (w, T, flag) 9 M1 + per-ray 8N")
    B["fine compact2"] = ("hbm", g(1 * M1 + 8 * M2), "GB/s", "flags M1 + ids 8 M2")
    B["fine feat"] = ("hbm", g(8 * M2 + 16 * M2 + 48 * uv["k0_level2"] + 4 * uv["sdf_level2"] + (448 + 256 + 16) * M2), "GB/s",
                      "ids + level-1 values 24 M2 + unique k0 voxels x 48 + unique sdf voxels x 4 + rows (X0 448 + XR tail "
                      "256 + normal 16) M2")
    B["fine composite"] = ("hbm", g(40 * M2 + 20 * M2 + 12 * N + 28 * N + 8 * N), "GB/s",
                           "logit, normal, w, idx 40 M2 + dlogit, d_w 20 M2 + target 12N + outputs 28N + offsets 8N")
    B["fine feat_bwd"] = ("hbm", g((448 + 96 + 256 + 256 + 24) * M2 + 96 * uv["k0_level2"] + 8 * uv["sdf_level2"] + 16 * M2), "GB/s",
                          "dX0 448 + taps of X0 96 + XR tail and its gradient 512 + ids 24 per sample, k0 scatter unique "
                          "voxels x 96 (RMW), sdf scatter unique voxels x 8, d_sdf1 / d_grad1 16 M2")
    B["fine weights1_bwd"] = ("hbm", g(17 * M1 + 4 * vis + 12 * N), "GB/s", "flag, d_w, w, alpha, T 17 M1 + d_alpha 4 visited + per-ray 12N")
    B["fine level1_bwd"] = ("hbm", g(9 * M1 + 28 * vis + 16 * M2 + 8 * uv["sdf_level1_visited"]), "GB/s",
                            "ids + flag 9 M1, (d_alpha, sdf, grad, step) 28 visited, level-2 gradients 16 M2, sdf scatter "
                            "unique voxels x 8 (RMW)")
    B["adam_dense_zero"] = ("hbm", g(32.0 * V), "GB/s", "p, g, m, v read + p, m, v, g written: 32 V")
    B["adam_bricks"] = ("hbm", g(bricks * 64 * 12 * 32.0), "GB/s", "touched 4^3 bricks x 64 voxels x 12 ch x 32 B")
    B["adam_upd_multi"] = ("hbm", g(28.0 * 436000), "GB/s", "16 MLP tensors, 0.436 M params x 28 B")
    B["weight_split"] = ("hbm", g(4.0 * 436000 + 16.0 * 436000), "GB/s", "weights read once, 4 planes written")
    B["smooth_tv err"] = ("hbm", g(24.0 * V * tv_frac), "GB/s", "G read, E written (3 ch) on TV steps")
    B["smooth_tv grad grid"] = ("hbm", g(16.0 * V * tv_frac), "GB/s", "sdf read, G written on TV steps")
    B["smooth_tv scatter"] = ("hbm", g(20.0 * V * tv_frac), "GB/s", "E read, g_sdf RMW on TV steps")
    B["total_variation_add_grad_v4"] = ("hbm", g(12.0 * V * tv_frac), "GB/s", "sdf read, g_sdf RMW: 12 V on TV steps")
    B["sdf_reg_fused"] = ("hbm", g(12.0 * V * tv_frac), "GB/s", "sdf read, g_sdf RMW: 12 V on TV steps (both regularisers in one pass)")
    return B


def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; fgs_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        # stdout carries exactly one JSON line: anything the communicator set-up prints ("NCCL version ...")
        # goes to stderr
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.all_reduce(torch.zeros(1, device=dev))      # communicators are created lazily
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    if args.mode == "render":
        return run_render(args, world, rank, dev)
    if args.stage in ("coarse", "geometry_searching"):
        if world > 1:
            raise SystemExit("bench.py --stage coarse / geometry_searching: these stages run on one GPU")
        return run_coarse(args, dev)
    from fgs_nerf_b200 import _lib, engine, scene

    N = args.rays
    sc = scene.make_scene(G=args.grid)
    eng = engine.FineStepEngine(sc, N, dev, world_size=world, rank=rank, cap_m2_per_ray=args.cap_m2_per_ray,
                                voxel_exchange=(args.exchange == "voxels"))
    eng.set_nonempty_mask()
    if args.dw_sm_reserve >= 0:
        eng.dw_sm_reserve = args.dw_sm_reserve
    eng.use_graph = bool(args.graph) and world == 1
    if args.freeze:
        eng.lr = {k: 0.0 for k in eng.lr}
    cam = camera(args)
    nb = 4
    host = []
    for b in range(nb):
        rays = scene.make_rays(N=N, offset=(rank * nb + b) * N, **cam)
        host.append(torch.stack(rays).pin_memory())       # one pinned [4, N, 3] block per batch: rays_o, rays_d, viewdirs, target
    resident = [[t.contiguous() for t in h.to(dev)] for h in host]
    n_glob = N * world
    feed = lambda ro, rd, vd, tgt, i: eng.train_step(ro, rd, vd, tgt, args.global_step + i, n_rays_global=n_glob,
                                                     has_last_ray=(rank == world - 1))
    pipe = E2EPipe(torch, host, dev, 5, feed)
    pipe_serial = E2EPipe(torch, host, dev, 5, feed, serial=True)

    def step_resident(i):
        ro, rd, vd, tgt = resident[i % nb]
        eng.train_step(ro, rd, vd, tgt, args.global_step + i, n_rays_global=n_glob, has_last_ray=(rank == world - 1))

    def step_e2e(i):
        pipe.step(i, eng)

    def step_e2e_serial(i):
        pipe_serial.step(i, eng)

    def timed(fn, k, i0):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(k):
            fn(i0 + i)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.barrier()
        return float(ms)

    it = 0
    for _ in range(args.warmup):
        step_resident(it); it += 1
    torch.cuda.synchronize()
    cnt0 = eng.counters()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    _lib.lib.fgs_reset_launch_count()
    ms = timed(step_resident, args.steps, it); it += args.steps
    launches = int(_lib.lib.fgs_launch_count())
    clocks = sampler.stop() if sampler else None
    cnt1 = eng.counters()
    for _ in range(2):
        step_e2e(it); it += 1
    ms_e2e = timed(step_e2e, args.steps, it); it += args.steps
    cnt2 = eng.counters()
    loss_host = pipe.loss().clone()
    step_e2e_serial(it); it += 1
    ms_e2e_serial = timed(step_e2e_serial, args.steps, it); it += args.steps

    # ---- per-launch timing of one more K steps (device events between launches) ----
    prof = {}
    torch.cuda.synchronize()                 # every rank runs these steps (they contain the all-reduce);
    if rank == 0:                            # only rank 0 records the per-launch events
        _lib.prof_begin()
    tv_steps = 0
    for i in range(args.steps):
        tv_steps += int(eng.is_tv_step(args.global_step + it + i))
        step_resident(it + i)
    if rank == 0:
        rows = _lib.prof_end()
        for name, t in rows:
            if name == "fork":          # stream-segment marker, not a launch (its interval spans the previous step)
                continue
            a = prof.setdefault(name, [0.0, 0])
            a[0] += t; a[1] += 1
    torch.cuda.synchronize()
    cnt3 = eng.counters()                    # the sample counts of the per-launch region (its last step)
    it += args.steps
    eng.check_overflow()
    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    K = args.steps
    V = args.grid ** 3
    # k0 voxels / 4^3 bricks that receive a gradient in a step and the unique voxels the gathers touch (one extra
    # forward + backward, untimed; the optimiser clears the gradient it consumes, so it cannot be read back after a
    # full step)
    ro, rd, vd, tgt = resident[(it - 1) % nb]
    eng.forward(ro, rd, vd, args.global_step + it - 1)
    eng.zero_grad()
    eng.loss_backward(tgt, n_glob, rank == world - 1)
    torch.cuda.synchronize()
    touched = int((eng.g_k0 != 0).any(dim=1).sum().item())
    bricks = int(eng.ws["k0_brick_flags"].sum().item())
    uv = unique_voxel_counts(eng, ro, rd)
    cnt_model = eng.counters()
    eng._grads_clean = False
    eng.zero_grad()
    step_ms_prof = sum(v[0] for v in prof.values()) / K
    kern = []
    for name, (t, n) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
        kern.append({"name": name, "ms_per_step": round(t / K, 5), "launches_per_step": round(n / K, 2),
                     "share": round(t / K / step_ms_prof, 4)})
    alg = byte_models(cnt_model, uv, N, V, tv_steps / K, bricks)
    traffic = ncu_traffic()
    for k in kern:
        if k["name"] in alg:
            bound, work, unit, model = alg[k["name"]]
            ach = work / (k["ms_per_step"] * 1e-3) if k["ms_per_step"] > 0 else 0.0
            k.update(bound=bound, achieved=round(ach, 2), unit=unit, model=model,
                     frac=round(ach / (pk["tensor"] if bound == "tensor" else pk["hbm"]), 4))
        if k["name"] in traffic:
            k["ncu_dram_bytes_per_launch"] = traffic[k["name"]]
    top = kern[0]
    tr = traffic.get(top["name"])
    roof = {"kernel": top["name"], "bound": top.get("bound", "hbm"), "achieved": top.get("achieved"),
            "peak": pk["tensor"] if top.get("bound") == "tensor" else pk["hbm"],
            "unit": top.get("unit", "GB/s"), "frac": top.get("frac"),
            "traffic": tr.get("bytes") if isinstance(tr, dict) else tr,
            "traffic_source": (tr.get("source") if isinstance(tr, dict) else None) or "profiles/r02_ncu_traffic.json",
            "peak_source": pk["src"], "model": top.get("model"),
            "launches_per_step": top["launches_per_step"],
            "avg_launch_ms": round(top["ms_per_step"] / max(top["launches_per_step"], 1e-9), 5),
            "share_of_step": top["share"],
            "M2_of_this_region": cnt_model["M2"],
            "note": "MLP GEMMs: algorithmic fp32 FLOPs of the launch (2 x MACs x M2 of THIS region) against the measured "
                    "bf16 dense peak; the forward runs 3xTF32 (3 MMAs per product at half the bf16 rate: ceiling 1/6 of "
                    "the peak), the backward kernels 3xBF16 (ceiling 1/3)"}
    spr = lambda c: {"total_len": c["total_len"] / N, "M0": c["M0"] / N, "M1": c["M1"] / N, "Ma": c["MA"] / N, "M2": c["M2"] / N}
    line = {"metric": METRIC, "value": n_glob * K / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": args.warmup, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": dict(workload(args), parallelism=f"ray-sharded dp{world}", global_rays_per_step=n_glob,
                           samples_per_ray=dict(spr(cnt1), M2_at_start=cnt0["M2"] / N),
                           M2_per_region={"value": cnt1["M2"], "e2e": cnt2["M2"], "per_launch": cnt3["M2"],
                                          "byte_model": cnt_model["M2"]},
                           k0_voxels_touched=touched, k0_bricks_touched=bricks, unique_voxels=uv),
            "clocks": clocks,
            "e2e": {"value": n_glob * K / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e / K,
                    "h2d_bytes_per_step": int(host[0].numel() * 4),
                    "d2h_bytes_per_step": int(loss_host.numel() * 8 + 8 * 4),
                    "mode": "pipelined host loop: inputs of step i+1 uploaded on a copy stream while step i runs, step i's "
                            "loss + counters read by the host after step i+1 is enqueued (every step's copies inside the "
                            "timed region)",
                    "serial_value": n_glob * K / (ms_e2e_serial * 1e-3), "serial_ms_per_step": ms_e2e_serial / K,
                    "serial_mode": "copy, step, D2H, synchronise -- nothing overlapped"},
            "gpu_launches": launches, "roofline": roof, "kernels": kern,
            "kernel_sum_ms_per_step": round(step_ms_prof, 5),
            "loss_last": [round(float(x), 6) for x in loss_host.tolist()]}
    del eng
    torch.cuda.empty_cache()
    if world == 1 and not args.no_gpu_oracle:
        try:
            line["gpu_oracle"] = gpu_oracle(args, dev)
        except Exception as e:                      # a baseline leg must never take the product's line down
            line["gpu_oracle"] = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args)
    if world > 1:
        dist.destroy_process_group()
    print(json.dumps(line), flush=True)


def run_coarse(args, dev):
    """one coarse-stage training iteration (forward_coarse + loss incl. weight_rgbper + backward + both
    TV terms + masked Adam) on one GPU: config/shiny_blender.py `coarse_model` / `coarse_train`, ~1.5 M voxels."""
    import torch
    from fgs_nerf_b200 import _lib, coarse, scene
    N = args.rays
    geometry = args.stage == "geometry_searching"
    if geometry:
        # the first stage of a run: no MaskCache, k0_dim 6, refnet 72-128-128-3, the voxel_inc box 60 % through its schedule
        sc = scene.make_scene(G=args.grid, with_mlps=False, cfg=dict(scene.FINE_MODEL, k0_dim=6))
        eng = coarse.CoarseStepEngine(sc, coarse.make_refnet(coarse.GEOMETRY_MODEL), N, dev, cfg=coarse.GEOMETRY_MODEL,
                                      train_cfg=coarse.GEOMETRY_TRAIN)
        w = 0.6
        eng.set_inc_mask([0.2 - w * 0.2] * 3, [0.8 + w * 0.2] * 3)
    else:
        sc = scene.make_scene(G=args.grid, with_mlps=False)
        eng = coarse.CoarseStepEngine(sc, coarse.make_refnet(), N, dev)
    eng.set_nonempty_mask()
    if args.freeze:
        eng.lr = {k: 0.0 for k in eng.lr}
    cam = camera(args)
    nb = 4
    host = [torch.stack(scene.make_rays(N=N, offset=b * N, **cam)).pin_memory() for b in range(nb)]
    resident = [[t.contiguous() for t in h.to(dev)] for h in host]
    pipe = E2EPipe(torch, host, dev, 7, lambda ro, rd, vd, tgt, i: eng.train_step(ro, rd, vd, tgt, args.global_step + i))

    def step_resident(i):
        eng.train_step(*resident[i % nb], args.global_step + i)

    def step_e2e(i):
        pipe.step(i, eng)

    def timed(fn, k, i0):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(k):
            fn(i0 + i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    it = 0
    for _ in range(args.warmup):
        step_resident(it); it += 1
    sampler = ClockSampler(dev.index)
    sampler.start()
    _lib.lib.fgs_reset_launch_count()
    K = args.steps
    ms = timed(step_resident, K, it); it += K
    launches = int(_lib.lib.fgs_launch_count())
    clocks = sampler.stop()
    cnt = eng.counters()
    for _ in range(2):
        step_e2e(it); it += 1
    ms_e2e = timed(step_e2e, K, it); it += K
    loss_host = pipe.loss().clone()
    torch.cuda.synchronize()
    _lib.prof_begin()
    for i in range(K):
        step_resident(it + i)
    prof = {}
    for name, t in _lib.prof_end():
        a = prof.setdefault(name, [0.0, 0])
        a[0] += t; a[1] += 1
    eng.check_overflow()
    tot = sum(v[0] for v in prof.values()) / K
    kern = [{"name": n, "ms_per_step": round(t / K, 5), "launches_per_step": round(c / K, 2), "share": round(t / K / tot, 4)}
            for n, (t, c) in sorted(prof.items(), key=lambda kv: -kv[1][0])]
    V = args.grid ** 3
    line = {"metric": METRIC, "value": N * K / (ms * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": K, "warmup": args.warmup,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32",
            "data": "synthetic",
            "config": {"workload": (f"shiny-blender GEOMETRY_SEARCHING-stage training step (forward_coarse without a MaskCache, "
                                    f"voxel_inc box at 60 % of its schedule, loss incl. weight_rgbper + bwd + smooth-grad TV + "
                                    f"unmasked autograd SDF TV every step + masked Adam), {args.grid}^3 SDF + 6-ch k0 grid, 5^3 "
                                    f"Gaussian SDF smoothing per forward, refnet 72-128-128-3, {N} rays/step, global_step "
                                    f"{args.global_step}+") if geometry else
                                   (f"shiny-blender COARSE-stage training step (forward_coarse + loss incl. weight_rgbper + bwd + "
                                    f"smooth-grad TV + autograd SDF TV every step + masked Adam), {args.grid}^3 SDF + 12-ch k0 "
                                    f"grid, 5^3 Gaussian SDF smoothing per forward, refnet 90-192-192-3, {N} rays/step, "
                                    f"global_step {args.global_step}+"),
                       "stage": args.stage, "grid": args.grid, "rays_per_gpu": N, "frozen_trajectory": bool(args.freeze),
                       "l2_policy": f"k0 + grad + Adam state = 4 x {48 * V / 1e6:.0f} MB, larger than the 126 MB L2",
                       "samples_per_ray": {"total_len": cnt["total_len"] / N, "M1": cnt["M1"] / N, "M2": cnt["M2"] / N}},
            "clocks": clocks,
            "e2e": {"value": N * K / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e / K,
                    "h2d_bytes_per_step": int(host[0].numel() * 4), "d2h_bytes_per_step": 7 * 8 + 8 * 4},
            "gpu_launches": launches, "kernels": kern, "kernel_sum_ms_per_step": round(tot, 5),
            "loss_last": [round(float(x), 6) for x in loss_host.tolist()]}
    print(json.dumps(line), flush=True)


def run_render(args, world, rank, dev):
    """BASELINE config 5: full-image inference, chunks dealt round-robin to the ranks (model/utils.py:296-299)."""
    import torch
    import torch.distributed as dist
    from fgs_nerf_b200 import _lib, engine, scene
    H, W = args.image_h, args.image_w
    R = H * W
    sc = scene.make_scene(G=args.grid)
    eng = engine.FineStepEngine(sc, args.rays, dev, cap_m2_per_ray=args.cap_m2_per_ray, cap_m1_per_ray=args.cap_m1_per_ray)
    # rays of a whole image = the first H*W pixels of a seeded permutation over 40 views (every chunk mixes views:
    # the same per-chunk statistics as a single view would give on average, no view-dependent imbalance between ranks)
    cam = camera(args)
    cam.update(H=H, W=W)
    views = [[t.to(dev) for t in scene.make_rays(N=R, offset=v * R, **cam)[:3]] for v in range(2)]
    host = [[t.cpu().pin_memory() for t in views[0]]]

    def image(i):
        ro, rd, vd = views[i % len(views)]
        return eng.render(ro, rd, vd, global_step=args.global_step, near=2.0, bg=1.0, render_depth=True, render_grad=True,
                          world_size=world, rank=rank)

    def image_e2e(i):
        rays = [t.to(dev, non_blocking=True) for t in host[0]]
        out = eng.render(*rays, global_step=args.global_step, near=2.0, bg=1.0, render_depth=True, render_grad=True,
                         world_size=world, rank=rank)
        return [out[k].cpu() for k in ("rgb_marched", "depth", "normal_marched")]

    def timed(fn, k):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(k):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.barrier()
        return float(ms)

    for i in range(max(args.warmup, 1)):
        image(i)
    torch.cuda.synchronize()
    cnt = eng.counters()
    sampler = ClockSampler(dev.index) if rank == 0 else None
    _lib.lib.fgs_reset_launch_count()
    K = args.steps
    ms = timed(image, K)
    launches = int(_lib.lib.fgs_launch_count())
    clocks = sampler.stop() if sampler else None
    image_e2e(0)
    ms_e2e = timed(image_e2e, K)
    if rank == 0:
        n = args.rays
        line = {"metric": "render_rays_per_sec", "value": R * K / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K,
                "warmup": max(args.warmup, 1), "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
                "config": dict(workload(args), parallelism=f"chunk-sharded dp{world}",
                               samples_per_ray_last_chunk={"total_len": cnt["total_len"] / n, "M1": cnt["M1"] / n, "M2": cnt["M2"] / n}),
                "clocks": clocks,
                "e2e": {"value": R * K / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e / K,
                        "h2d_bytes_per_step": 36 * R, "d2h_bytes_per_step": 28 * R},
                "gpu_launches": launches}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="train", choices=["train", "render"])
    ap.add_argument("--stage", default="fine", choices=["fine", "coarse", "geometry_searching"],
                    help="coarse / geometry_searching: one iteration of that stage (SURVEY 8f row 1; use --grid 116 "
                         "--global-step 3000 / --grid 120 --global-step 600)")
    ap.add_argument("--grid", type=int, default=160)
    ap.add_argument("--rays", type=int, default=8192, help="rays per GPU per step (train) / chunk size (render)")
    ap.add_argument("--camera", default="blender", choices=["blender", "dtu"])
    ap.add_argument("--image-h", type=int, default=800)
    ap.add_argument("--image-w", type=int, default=800)
    ap.add_argument("--global-step", type=int, default=5000)
    ap.add_argument("--freeze", type=int, default=1, help="1: learning rates 0 (stationary sample counts); 0: train for real")
    ap.add_argument("--ref-rays", type=int, default=2048, help="rays per step of the CPU reference arm / baseline")
    ap.add_argument("--cap-m2-per-ray", type=int, default=32)
    ap.add_argument("--cap-m1-per-ray", type=int, default=192)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-oracle", action="store_true")
    ap.add_argument("--dw-sm-reserve", type=int, default=-1,
                    help="N > 1: SMs the weight-gradient kernel leaves to the overlapped exchange (-1: the engine's default)")
    ap.add_argument("--graph", type=int, default=1, help="1 (default): the single-GPU iteration is launched as one CUDA graph")
    ap.add_argument("--exchange", default="bricks", choices=["bricks", "voxels"],
                    help="granularity of the data-parallel k0 gradient exchange (N > 1)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours" and args.mode == "train":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
