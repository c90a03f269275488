#!/usr/bin/env python
"""bench.py -- train rays/s of the fused fine-stage step (fwd + bwd + TV + masked Adam).

    python bench.py --gpus N --steps K --warmup W            (our arm; torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K --warmup W   (reference's CPU path)

Workload (BASELINE.json configs[1]): shiny-blender fine stage, 160^3 SDF + 12-channel feature grid,
8192 rays per GPU per step, <= 256 post-mask samples per ray, synthetic scene of SURVEY.md 8d
(random-init grids + MLPs, blender pinhole rays, random targets), global_step 5000 onwards
(TV regularisers on every 3rd step, as config/shiny_blender.py:180-219 schedules them).

One JSON line on stdout (rank 0):
  value      rays/s, inputs resident in HBM, K steps timed with CUDA events, max over ranks
  e2e        same metric through the public host API: rays + targets start in pinned HOST memory,
             H2D copies and a D2H read of the step's loss + sample counters inside the timed region
  roofline   the dominant kernel of the step, timed per launch with CUDA events (fgs_prof_*)
  cpu_baseline  the oracle (torch CPU restatement of the reference step) on a bounded ray sample
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


def workload(args):
    return {"workload": f"shiny-blender fine-stage training step (fwd+bwd+TV every 3rd+masked Adam), "
                        f"{args.grid}^3 SDF + 12-ch k0 grid, {args.rays} rays/GPU/step, stepsize 0.5, "
                        f"global_step {args.global_step}+",
            "grid": args.grid, "rays_per_gpu": args.rays, "global_step": args.global_step,
            "l2_policy": "working set (k0 + grad + Adam state = 4 x 197 MB at 160^3) larger than the 126 MB L2"}


# ---------------------------------------------------------------------------------------------
def run_reference(args):
    """Reference arm: the reference's own CPU path for this step = oracle/nerf_ref.py (the torch
    restatement pinned against model/nerf.py) + oracle/fgs_oracle.c for the three native ops, on
    all host threads, a bounded ray sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from fgs_nerf_b200 import scene
    from oracle import native_ref, nerf_ref as R
    native_ref.build()
    n = args.ref_rays
    sc = scene.make_scene(G=args.grid)
    m = R.FineModel(sc)
    m.set_nonempty_mask()
    nat = R.CpuNative()
    opt = R.RefOptimizer(m, scene.FINE_TRAIN, nat)
    batches = [scene.make_rays(N=n, offset=b * n) for b in range(2)]
    times = []
    for i in range(args.warmup + args.steps):
        ro, rd, vd, tgt = batches[i % len(batches)]
        t0 = time.perf_counter()
        R.train_step(m, nat, opt, ro, rd, vd, tgt, args.global_step + i, scene.FINE_TRAIN, near=2.0, stepsize=0.5, bg=1.0)
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    v = n * len(times) / total
    sample = f"{n} rays/step of the same 160^3 workload ({len(times)} steps, {args.warmup} warm-up)"
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic", "config": workload(args),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_baseline(args):
    """Oracle step on a bounded sample, timed on the host cores (rank 0, N = 1 only)."""
    import torch
    from fgs_nerf_b200 import scene
    from oracle import native_ref, nerf_ref as R
    native_ref.build()
    n = 2 * args.ref_rays
    sc = scene.make_scene(G=args.grid)
    m = R.FineModel(sc)
    m.set_nonempty_mask()
    nat = R.CpuNative()
    opt = R.RefOptimizer(m, scene.FINE_TRAIN, nat)
    ro, rd, vd, tgt = scene.make_rays(N=n)
    times = []
    for i in range(3):
        t0 = time.perf_counter()
        R.train_step(m, nat, opt, ro, rd, vd, tgt, args.global_step + i, scene.FINE_TRAIN, near=2.0, stepsize=0.5, bg=1.0)
        times.append(time.perf_counter() - t0)
    t = min(times[1:])
    return {"value": n / t, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{n} rays of the same workload, best of 2 steps after 1 warm-up ({t:.2f} s/step)"}


# ---------------------------------------------------------------------------------------------
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
    from fgs_nerf_b200 import _lib, engine, scene

    N = args.rays
    sc = scene.make_scene(G=args.grid)
    eng = engine.FineStepEngine(sc, N, dev, world_size=world, rank=rank, cap_m2_per_ray=args.cap_m2_per_ray,
                                voxel_exchange=(args.exchange == "voxels"))
    eng.set_nonempty_mask()
    nb = 4
    host = []
    for b in range(nb):
        rays = scene.make_rays(N=N, offset=(rank * nb + b) * N)
        host.append([t.pin_memory() for t in rays])
    resident = [[t.to(dev) for t in h] for h in host]
    staging = [torch.empty_like(t, device=dev) for t in host[0]]
    loss_host = torch.empty(5, dtype=torch.float64).pin_memory()
    cnt_host = torch.empty(8, dtype=torch.int32).pin_memory()
    n_glob = N * world

    def step_resident(i):
        ro, rd, vd, tgt = resident[i % nb]
        eng.train_step(ro, rd, vd, tgt, args.global_step + i, n_rays_global=n_glob, has_last_ray=(rank == world - 1))

    def step_e2e(i):
        for s, h in zip(staging, host[i % nb]):
            s.copy_(h, non_blocking=True)
        eng.train_step(*staging, args.global_step + i, n_rays_global=n_glob, has_last_ray=(rank == world - 1))
        loss_host.copy_(eng.ws["loss"], non_blocking=True)
        cnt_host.copy_(eng.ws["counters"], non_blocking=True)
        torch.cuda.current_stream().synchronize()       # the caller reads the step's loss

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

    # ---- per-launch timing of one more K steps (device events between launches) ----
    prof = {}
    torch.cuda.synchronize()                 # every rank runs these steps (they contain the all-reduce);
    if rank == 0:                            # only rank 0 records the per-launch events
        _lib.prof_begin()
    for i in range(args.steps):
        step_resident(it + i)
    if rank == 0:
        rows = _lib.prof_end()
        for name, t in rows:
            if name == "fork":          # stream-segment marker, not a launch (its interval spans the previous step)
                continue
            a = prof.setdefault(name, [0.0, 0])
            a[0] += t; a[1] += 1
    torch.cuda.synchronize()
    it += args.steps
    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    K = args.steps
    M1, M2 = cnt1["M1"], cnt1["M2"]
    V = args.grid ** 3
    # k0 voxels / 4^3 bricks that receive a gradient in a step (one extra forward + backward, untimed;
    # the optimiser clears the gradient it consumes, so it cannot be read back after a full step)
    ro, rd, vd, tgt = resident[0]
    eng.forward(ro, rd, vd, args.global_step + it)
    eng.zero_grad()
    eng.loss_backward(tgt, n_glob, rank == world - 1)
    torch.cuda.synchronize()
    touched = int((eng.g_k0 != 0).any(dim=1).sum().item())
    bricks = int(eng.ws["k0_brick_flags"].sum().item())
    eng._grads_clean = False
    eng.zero_grad()
    tv_steps = sum(1 for i in range(K) if eng.is_tv_step(args.global_step + args.warmup + i))
    step_ms_prof = sum(v[0] for v in prof.values()) / K
    kern = []
    for name, (t, n) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
        kern.append({"name": name, "ms_per_step": round(t / K, 5), "launches_per_step": round(n / K, 2),
                     "share": round(t / K / step_ms_prof, 4)})
    # algorithmic work per step of the kernels that have a simple model ("byte model")
    # MLP GEMMs: forward and backward-data (gemm_w: X*W^T, dZ*W) and backward-weight (gemm_tc: dZ^T*X);
    # algorithmic FLOPs = 2 * MACs per product, fp32 semantics (the 3xTF32 split issues 3x that on the
    # tensor cores; the roofline below counts the algorithmic FLOPs only)
    mac = 2.0 * MLP_MACS * M2
    V_t = touched
    alg = {"gemm_w": ("tensor", 2 * mac / 1e12, "TFLOP/s"),
           "mlp_chain": ("tensor", 2 * mac / 1e12, "TFLOP/s"),      # forward chain + backward-data chain
           "gemm_tn": ("tensor", mac / 1e12, "TFLOP/s"),
           "gemm_tc": ("tensor", mac / 1e12, "TFLOP/s"),
           "gemm": ("tensor", 3 * mac / 1e12, "TFLOP/s"),
           "adam_dense_zero": ("hbm", 32.0 * V / 1e9, "GB/s"),                          # p,g,m,v read; p,m,v,g written
           "adam_bricks": ("hbm", 12.0 * V_t * 32.0 / 1e9, "GB/s"),                    # touched voxels x 12 ch x 32 B
           "adam_upd": ("hbm", (28.0 * V + 4.0 * 12 * V + 24.0 * 12 * touched) / 1e9, "GB/s"),
           "memset": ("hbm", (4.0 * V + 48.0 * V) / 1e9, "GB/s"),
           "smooth_tv err": ("hbm", 24.0 * V / 1e9 * tv_steps / K, "GB/s"),               # G read, E written (3 ch)
           "smooth_tv grad grid": ("hbm", 16.0 * V / 1e9 * tv_steps / K, "GB/s"),         # sdf read, G written
           "smooth_tv scatter": ("hbm", 20.0 * V / 1e9 * tv_steps / K, "GB/s"),           # E read, g_sdf RMW
           "total_variation_add_grad_v4": ("hbm", 12.0 * V / 1e9 * tv_steps / K, "GB/s")}
    for k in kern:
        if k["name"] in alg:
            bound, work, unit = alg[k["name"]]
            ach = work / (k["ms_per_step"] * 1e-3) if k["ms_per_step"] > 0 else 0.0
            k.update(bound=bound, achieved=round(ach, 2), unit=unit,
                     frac=round(ach / (pk["tensor"] if bound == "tensor" else pk["hbm"]), 4))
    top = kern[0]
    # DRAM traffic of the dominant kernel per launch from the committed ncu --set full capture of this command
    # (profiles/r01_ncu_full_mlp_chain2_dw.txt: dram__bytes_read.sum + dram__bytes_write.sum, taken at M2 = 42.5 K)
    ncu_traffic = {"mlp_chain": 37.42e6 + 263.94e6, "mlp_dw": 1051.6e6 + 7.32e6}
    roof = {"kernel": top["name"], "bound": top.get("bound", "hbm"), "achieved": top.get("achieved"),
            "peak": pk["tensor"] if top.get("bound") == "tensor" else pk["hbm"],
            "unit": top.get("unit", "GB/s"), "frac": top.get("frac"), "traffic": ncu_traffic.get(top["name"]),
            "traffic_source": "profiles/r01_ncu_full_mlp_chain2_dw.txt (bytes per launch at M2 = 42.5 K samples)",
            "peak_source": pk["src"],
            "launches_per_step": top["launches_per_step"], "avg_launch_ms": round(top["ms_per_step"] / max(top["launches_per_step"], 1e-9), 5),
            "share_of_step": top["share"],
            # the same achieved rate against what this formulation can reach at best: kind::tf32 issues at half the
            # bf16 rate and 3xTF32 spends three MMAs per product
            "frac_of_3xtf32_ceiling": (round(top["achieved"] / (pk["tensor"] / 6.0), 4)
                                       if top.get("bound") == "tensor" and top.get("achieved") else None),
            "note": "3xTF32 tcgen05 GEMMs of the colour MLPs: algorithmic fp32 FLOPs against the measured bf16 tensor "
                    "peak (kind::tf32 runs at half the bf16 rate and the split issues 3 MMAs per product, so the "
                    "ceiling of this formulation is 1/6 of that peak)"}
    line = {"metric": METRIC, "value": n_glob * K / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": args.warmup, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": dict(workload(args), parallelism=f"ray-sharded dp{world}", global_rays_per_step=n_glob,
                           samples_per_ray={"total_len": cnt1["total_len"] / N, "M0": cnt1["M0"] / N, "M1": M1 / N,
                                            "Ma": cnt1["MA"] / N, "M2": M2 / N, "M2_at_start": cnt0["M2"] / N},
                           k0_voxels_touched=touched, k0_bricks_touched=bricks),
            "clocks": clocks,
            "e2e": {"value": n_glob * K / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e / K,
                    "h2d_bytes_per_step": int(sum(t.numel() * 4 for t in host[0])),
                    "d2h_bytes_per_step": int(loss_host.numel() * 8 + cnt_host.numel() * 4)},
            "gpu_launches": launches, "roofline": roof, "kernels": kern,
            "loss_last": [round(float(x), 6) for x in loss_host.tolist()]}
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args)
    if world > 1:
        dist.destroy_process_group()
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", type=int, default=160)
    ap.add_argument("--rays", type=int, default=8192)
    ap.add_argument("--global-step", type=int, default=5000)
    ap.add_argument("--ref-rays", type=int, default=2048, help="rays per step of the CPU reference arm / baseline")
    ap.add_argument("--cap-m2-per-ray", type=int, default=16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="bricks", choices=["bricks", "voxels"],
                    help="granularity of the data-parallel k0 gradient exchange (N > 1)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
