#!/usr/bin/env python
"""CPU numerics probe for the next MLP kernel generation: which operand split keeps the 1e-4 parity bar?

The MLPs run today as 3xTF32 (hi/lo TF32 planes, 3 tcgen05.mma.kind::tf32 per product) and are bound by
the shared-memory operand bandwidth of fp32-sized operands.  kind::f16 (bf16 operands) issues at twice the TF32 rate
and reads half the operand bytes, so a bf16 split with 3 MMAs per product (hh + hm + mh) would halve both limits.
This script emulates the splits on the CPU (operands rounded to bf16 / tf32 planes, products and sums in fp32 like
the tensor core's fp32 accumulator; the TMEM truncation is not modelled) on the real network shapes -- rgbnet 106 ->
256 x3 -> 256, refnet 307 -> 256 x3 -> 3, nn.Linear init -- forward, backward-data and weight gradients, against fp64.

    python tools/split_numerics_probe.py [n_samples]
"""
import json
import sys

import torch

torch.manual_seed(0)


def rn_bf16(x):
    return x.to(torch.bfloat16).to(torch.float32)


def rn_tf32(x):
    # round-to-nearest-even to 10 explicit mantissa bits
    i = x.contiguous().view(torch.int32)
    r = ((i >> 13) & 1) + 0x0FFF
    return ((i + r) & ~0x1FFF).view(torch.float32)


def planes(x, rnd, n):
    out, rem = [], x
    for _ in range(n):
        p = rnd(rem)
        out.append(p)
        rem = rem - p
    return out


def mm_split(a, b, rnd, terms):
    """a @ b with both operands split into planes; `terms` = list of (i, j) plane products that are issued."""
    n = max(max(t) for t in terms) + 1
    pa, pb = planes(a, rnd, n), planes(b, rnd, n)
    acc = torch.zeros(a.shape[0], b.shape[1])
    for i, j in sorted(terms, key=lambda t: -(t[0] + t[1])):          # small terms first
        acc = acc + pa[i] @ pb[j]
    return acc


SCHEMES = {
    "fp32 (torch CPU)": None,
    "3xTF32  (hh+hl+lh)          [today]": (rn_tf32, [(0, 0), (0, 1), (1, 0)]),
    "1xTF32": (rn_tf32, [(0, 0)]),
    "3xBF16  (hh+hm+mh)": (rn_bf16, [(0, 0), (0, 1), (1, 0)]),
    "4xBF16  (hh+hm+mh+mm)": (rn_bf16, [(0, 0), (0, 1), (1, 0), (1, 1)]),
    "6xBF16  (all i+j<=2)": (rn_bf16, [(0, 0), (0, 1), (1, 0), (1, 1), (0, 2), (2, 0)]),
}


def lin_init(i, o):
    b = 1.0 / i ** 0.5
    return (torch.rand(o, i) * 2 - 1) * b, (torch.rand(o) * 2 - 1) * b


def run(M):
    dims = [(106, 256), (256, 256), (256, 256), (256, 256), (307, 256), (256, 256), (256, 256), (256, 3)]
    Ws = [lin_init(i, o) for i, o in dims]
    x0 = torch.randn(M, 106) * 0.5
    emb = torch.randn(M, 51) * 0.7
    dlogit = torch.randn(M, 3) * 1e-3

    def chain(mm, dt, use_masks=None):
        """forward, backward-data, weight gradients with matmul `mm`; returns logits, dX0, [dW], relu masks.
        use_masks: take the ReLU branches of another run (the parity tests compare gradients at the oracle's branch
        assignment, DESIGN.md 5.1 'ReLU kinks')."""
        acts, masks = [], []
        h = x0.to(dt)
        for l, (w, b) in enumerate(Ws):
            if l == 4:
                h = torch.cat([h, emb.to(dt)], 1)
            acts.append(h)
            z = mm(h, w.to(dt).T) + b.to(dt)
            if l in (3, 7):
                h = z
                masks.append(None)
            else:
                masks.append(z > 0)
                h = torch.relu(z)
        logits = h
        own_masks = masks
        if use_masks is not None:
            masks = use_masks
        g = dlogit.to(dt)
        dWs = [None] * 8
        for l in range(7, -1, -1):
            if masks[l] is not None:
                g = g * masks[l].to(dt)
            dWs[l] = mm(g.T.contiguous(), acts[l])
            g = mm(g, Ws[l][0].to(dt))
            if l == 4:
                g = g[:, :256]
        return logits, g, dWs, own_masks

    ref = chain(lambda a, b: a @ b, torch.float64)

    def rel(a, b):
        return float((a.double() - b).abs().max() / b.abs().max())

    rows = []
    for name, sch in SCHEMES.items():
        mm = (lambda a, b: a @ b) if sch is None else (lambda a, b, s=sch: mm_split(a, b, s[0], s[1]))
        out = chain(mm, torch.float32)
        flips = sum(int((m != r).sum()) for m, r in zip(out[3], ref[3]) if m is not None)
        units = sum(m.numel() for m in ref[3] if m is not None)
        at_ref = chain(mm, torch.float32, use_masks=ref[3])
        rows.append(dict(scheme=name, logits=rel(out[0], ref[0]), relu_flips=flips, flip_fraction=flips / units,
                         dX0_own_branches=rel(out[1], ref[1]), dX0_at_ref_branches=rel(at_ref[1], ref[1]),
                         dW_at_ref_branches=max(rel(a, b) for a, b in zip(at_ref[2], ref[2]))))
        print(json.dumps(rows[-1]))
    return rows


if __name__ == "__main__":
    run(int(sys.argv[1]) if len(sys.argv) > 1 else 4096)
