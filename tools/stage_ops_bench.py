#!/usr/bin/env python
"""Roofline numbers for the between-step operators (include/fgs_b200_train.h), at the sizes
of the shiny-blender fine stage: the pg_scale step 160^3 -> 256^3 (config/shiny_blender.py:203-204), the stage
hand-off smoother on 256^3, a 512^3 field query and the per-iteration batch gather; plus the iteration rate of
FineTrainer (sampler + gather + step + schedules) beside a bare engine.train_step loop on a fixed batch.

    python tools/stage_ops_bench.py [--out profiles/rNN_stage_ops.json]

achieved = ALGORITHMIC bytes (each input and output element once) / CUDA-event time, best of 5 after 2 warm-ups,
against MEASURED_PEAKS.json's HBM copy bandwidth.  Each op is also timed against the library op the reference calls
(F.interpolate, nn.Conv3d, F.grid_sample on materialised points, 4 index selects) on the same device.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--steps", type=int, default=30)
    args = ap.parse_args()
    from fgs_nerf_b200 import engine, extract_geometry, scene, training
    dev = torch.device("cuda:0")
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    rows = []

    def row(name, nbytes, ms, ref_ms, note):
        gbs = nbytes / ms / 1e6
        rows.append(dict(op=name, algorithmic_bytes=int(nbytes), ms=round(ms, 4), achieved_gbs=round(gbs, 1),
                         peak_gbs=peak, frac=round(gbs / peak, 4), reference_op_ms=round(ref_ms, 4),
                         speedup_vs_reference_op=round(ref_ms / ms, 2), note=note))
        print(json.dumps(rows[-1]), flush=True)

    G, G2 = 160, 256
    k0 = torch.randn(1, 12, G, G, G, device=dev)
    k0cl = k0.contiguous(memory_format=torch.channels_last_3d)
    sdf = torch.randn(1, 1, G, G, G, device=dev)
    ms = timed(lambda: engine.grid_resize(k0cl, (G2, G2, G2)))
    ref = timed(lambda: F.interpolate(k0, size=(G2, G2, G2), mode="trilinear", align_corners=True))
    row("grid_resize k0 12ch 160^3->256^3 (channels-last)", 4 * 12 * (G ** 3 + G2 ** 3), ms, ref, "F.interpolate NCDHW")
    ms = timed(lambda: engine.grid_resize(sdf, (G2, G2, G2)))
    ref = timed(lambda: F.interpolate(sdf, size=(G2, G2, G2), mode="trilinear", align_corners=True))
    row("grid_resize sdf 160^3->256^3", 4 * (G ** 3 + G2 ** 3), ms, ref, "F.interpolate")
    del k0, k0cl

    big = torch.randn(1, 1, G2, G2, G2, device=dev)
    w = engine.gaussian_3dconv_weights(5, 1, dev)
    conv = torch.nn.Conv3d(1, 1, 5, stride=1, padding=2, padding_mode="replicate").to(dev)
    conv.weight.data = w[None, None].to(dev)
    conv.bias.data = torch.zeros(1, device=dev)
    ms = timed(lambda: engine.smooth_conv(big, 5, 1, weights=w))
    with torch.no_grad():
        ref = timed(lambda: conv(big))
    row("smooth_conv 5^3 on 256^3", 8 * G2 ** 3, ms, ref, "nn.Conv3d replicate padding (cuDNN)")

    res = 512
    bmin, bmax = torch.tensor([-1.0] * 3), torch.tensor([1.0] * 3)
    mn, mx = torch.tensor([-1.0] * 3), torch.tensor([1.0] * 3)
    ms = timed(lambda: extract_geometry.extract_sdf_field(big, mn, mx, bmin, bmax, res), reps=3, warm=1)

    def ref_field():
        neg = -big
        q = lambda pts: F.grid_sample(neg, (((pts.to(dev) - mn.to(dev)) / (mx.to(dev) - mn.to(dev))).flip((-1,)) * 2 - 1
                                            ).reshape(1, 1, 1, -1, 3), mode="bilinear", align_corners=True).reshape(-1)
        return extract_geometry.extract_fields(bmin.to(dev), bmax.to(dev), res, q, N=64)
    import time
    t0 = time.time()
    ref_field()
    ref = (time.time() - t0) * 1e3
    row("sdf field query 512^3 from 256^3", 4 * (res ** 3 + G2 ** 3), ms, ref,
        "reference-style 64^3 block loop: materialised points + F.grid_sample + one D2H (wall clock)")
    del big

    n_total, N = 4_000_000, 8192
    arrs = [torch.randn(n_total, 3, device=dev) for _ in range(4)]
    sel = torch.randint(0, n_total, (N,), device=dev)
    out = tuple(torch.empty(N, 3, device=dev) for _ in range(4))
    ms = timed(lambda: engine.gather_rays(*arrs, sel, out=out), reps=20)
    sel_cpu = sel.cpu()
    ref = timed(lambda: [a[sel_cpu] for a in arrs], reps=20)
    row("gather_rays 8192 of 4M", N * (8 + 4 * 24), ms, ref, "4 index selects with a CPU index tensor, as nerf_training.py:253-258")

    # iteration rate: the loop (sampler + gather + step + schedules) against the bare step on a resident batch
    sc = scene.make_scene(G=G, mask_G=120)
    ro, rd, vd, tgt = [t.to(dev) for t in scene.make_rays(N=n_total // 8)]
    eng = engine.FineStepEngine(sc, N, dev)
    eng.set_nonempty_mask()
    tr = training.FineTrainer(eng, ro, rd, vd, tgt, cfg_train=dict(scene.FINE_TRAIN, N_rand=N, lrate_decay=20), start=5000,
                              rng=np.random.RandomState(0))
    tr.run(5)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    tr.run(args.steps)
    b.record()
    b.synchronize()
    loop_ms = a.elapsed_time(b) / args.steps
    st = tr.stats()
    batch = [t[:N].contiguous() for t in (ro, rd, vd, tgt)]
    import time as _time

    def bare(native, first):
        eng.native_step = native
        for i in range(5):
            eng.train_step(*batch, first + i)
        torch.cuda.synchronize()
        a.record()
        t0 = _time.perf_counter()
        for i in range(args.steps):
            eng.train_step(*batch, first + 100 + i)
        cpu_ms = (_time.perf_counter() - t0) * 1e3 / args.steps      # host time to ENQUEUE an iteration
        b.record()
        b.synchronize()
        return a.elapsed_time(b) / args.steps, cpu_ms

    host_ms, host_cpu = bare(False, 5100)
    lib_ms, lib_cpu = bare(True, 5400)
    rows.append(dict(op="training iteration, 160^3, 8192 rays: FineTrainer loop / bare step sequenced by the library / by Python",
                     trainer_loop_ms_per_step=round(loop_ms, 4), trainer_loop_rays_per_s=round(N / loop_ms * 1e3),
                     library_sequenced_ms_per_step=round(lib_ms, 4), library_sequenced_host_enqueue_ms=round(lib_cpu, 4),
                     python_sequenced_ms_per_step=round(host_ms, 4), python_sequenced_host_enqueue_ms=round(host_cpu, 4),
                     psnr=round(st["psnr"], 3), steps=args.steps))
    print(json.dumps(rows[-1]), flush=True)
    if args.out:
        with open(args.out, "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
