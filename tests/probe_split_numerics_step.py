#!/usr/bin/env python
"""Whole-step numerics probe (CPU, not a test): what happens to the PARITY QUANTITIES -- rendered colours, the grid
gradients, the weight gradients -- when the MLP GEMMs use a cheaper operand split than today's 3xTF32?

Runs the oracle's fine-stage step (oracle/nerf_ref.py, the checker of the GPU parity tests) twice on the same synthetic
scene: once in plain fp32 and once with every Linear of rgbnet / refnet replaced by an emulated split GEMM (forward,
backward-data and weight gradient; operands rounded to bf16 / tf32 planes, fp32 accumulation), with the ReLU branches of
the fp32 run injected -- exactly how the GPU tests compare gradients ("ReLU kinks").  Prints the two
norms the tests bound by 1e-4.  Lives under tests/ because it imports the oracle.

    python tests/probe_split_numerics_step.py [G] [N]
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from oracle import nerf_ref as R  # noqa: E402
from split_numerics_probe import SCHEMES, mm_split  # noqa: E402


class SplitLinear(torch.autograd.Function):
    """Linear whose three GEMMs (forward, backward-data, weight gradient) use emulated operand splits; `bwd` may name
    a different split for the two backward products."""

    @staticmethod
    def forward(ctx, x, w, b, fwd, bwd):
        ctx.save_for_backward(x, w)
        ctx.s = bwd
        return mm_split(x, w.T.contiguous(), fwd[0], fwd[1]) + b

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        rnd, terms = ctx.s
        g = g.contiguous()
        return mm_split(g, w, rnd, terms), mm_split(g.T.contiguous(), x, rnd, terms), g.sum(0), None, None


def run(G, N):
    from fgs_nerf_b200 import scene
    sc = scene.make_scene(G=G, mask_G=min(120, G + G // 4))
    ro, rd, vd, tgt = scene.make_rays(N=N, H=800 if G > 64 else 100, W=800 if G > 64 else 100, n_views=40 if G > 64 else 8)
    cfg = dict(scene.FINE_TRAIN, tv_every=1)
    native = R.CpuNative()
    orig_mlp = R.mlp
    masks, units = [], [0]

    def rec_mlp(x, layers):
        for i, (w, b) in enumerate(layers):
            x = torch.nn.functional.linear(x, w, b)
            if i + 1 < len(layers):
                masks.append(x > 0)
                units[0] += x.numel()
                x = torch.relu(x)
        return x

    def grads(m):
        return {k: p.grad.detach().clone() for k, p in m.params()}

    R.mlp = rec_mlp
    m0 = R.FineModel(sc)
    m0.set_nonempty_mask()
    out0, _ = R.train_step(m0, native, None, ro, rd, vd, tgt, 5000, cfg, near=2.0, stepsize=0.5, bg=1.0)
    g0 = grads(m0)
    print(json.dumps(dict(G=G, N=N, M2=int(out0["weights"].numel()), relu_units=units[0])))

    def rel(a, b):
        a, b = a.detach().double(), b.detach().double()
        return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300)), float((a - b).norm() / b.norm().clamp_min(1e-300))

    t32 = SCHEMES["3xTF32  (hh+hl+lh)          [today]"]
    bf3 = SCHEMES["3xBF16  (hh+hm+mh)"]
    cases = [(name, sch, sch) for name, sch in SCHEMES.items() if sch is not None]
    cases.append(("forward 3xTF32, backward-data + weight gradients 3xBF16", t32, bf3))
    for name, sch, sch_bwd in cases:
        it = iter(masks)
        flips = [0]

        def split_mlp(x, layers, sch=sch, sch_bwd=sch_bwd):
            for i, (w, b) in enumerate(layers):
                x = SplitLinear.apply(x, w, b, sch, sch_bwd)
                if i + 1 < len(layers):
                    mk = next(it)
                    flips[0] += int(((x > 0) != mk).sum())
                    x = x * mk
            return x

        R.mlp = split_mlp
        m1 = R.FineModel(sc)
        m1.set_nonempty_mask()
        out1, _ = R.train_step(m1, native, None, ro, rd, vd, tgt, 5000, cfg, near=2.0, stepsize=0.5, bg=1.0)
        g1 = grads(m1)
        same_samples = out1["ray_id"].shape == out0["ray_id"].shape and bool((out1["ray_id"] == out0["ray_id"]).all())
        row = dict(scheme=name.strip(), same_sample_set=same_samples, relu_flips=flips[0])
        row["rgb_marched max|L2"] = rel(out1["rgb_marched"], out0["rgb_marched"])
        for k in ("sdf", "k0", "rgbnet.0.weight", "refnet.0.weight", "refnet.3.weight"):
            row["grad " + k + " max|L2"] = rel(g1[k], g0[k])
        print(json.dumps(row))
    R.mlp = orig_mlp


if __name__ == "__main__":
    run(int(sys.argv[1]) if len(sys.argv) > 1 else 48, int(sys.argv[2]) if len(sys.argv) > 2 else 192)
