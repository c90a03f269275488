"""Regenerate the measured-results blocks of DESIGN.md (section 10) and README.md from the bench lines under profiles/
(between the markers <!-- results:begin --> / <!-- results:end -->).   usage: python tools/fill_results.py"""
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")


def line(name):
    f = os.path.join(P, name)
    if not os.path.exists(f):
        return None
    return json.loads(open(f).read().strip().splitlines()[-1])


def M(v):
    return f"{v / 1e6:.2f} M" if v >= 1e5 else f"{v / 1e3:.1f} K"


def row(label, d, fname, per="", extra=""):
    if d is None:
        return f"| {label} | not measured | | | |"
    e = d.get("e2e") or {}
    return (f"| {label} | {M(d['value'])} | {d['ms_per_step']:.3f}{per} | {M(e['value']) if e.get('value') else '--'} | "
            f"`{fname}`{extra} |")


def main():
    c2, c3, c5, co = line("r02_bench_cfg2.json"), line("r02_bench_cfg3.json"), line("r02_bench_cfg5_n1.json"), line("r02_bench_coarse.json")
    rf, n2, n8 = line("r02_bench_reference.json"), line("r02_bench_cfg2_n2.json"), line("r02_bench_cfg2_n8.json")
    c4, c58 = line("r02_bench_cfg4_n8.json"), line("r02_bench_cfg5_n8.json")
    # the 8-GPU lines were taken one build earlier (before march_kernel / feat_kernel changed by ~4.5 us per step):
    # their ratios are formed against the single-GPU line of THAT build and visit
    c2s = line("r02_bench_cfg2_prev.json") or c2
    c5p = line("r02_bench_cfg5_n1_prev.json") or c5
    go, cb = c2.get("gpu_oracle") or {}, c2.get("cpu_baseline") or {}
    spr = c2["config"]["samples_per_ray"]
    t = ["| configuration | rays/s (`value`) | ms / step | end to end (`e2e`) | file |", "|---|---|---|---|---|"]
    t.append(row(f"2: 160^3, 8192 rays, 1 GPU (headline; M2/ray {spr['M2']:.2f})", c2, "r02_bench_cfg2.json"))
    if go.get("value"):
        t.append(f"| the same step by the reference's arithmetic on the same B200 (`gpu_oracle`: torch CUDA ops + the reference's "
                 f"compiled kernels) | {M(go['value'])} | {go['ms_per_step']:.2f} | -- | same line: ours is "
                 f"{c2['value'] / go['value']:.1f}x |")
    if cb.get("value"):
        t.append(f"| the same step on the box's {cb['cores']} host cores (`cpu_baseline`, 4096-ray sample; `--impl reference`: "
                 f"2048-ray steps) | {M(cb['value'])} / {M(rf['value']) if rf else '--'} | -- | -- | same line / `r02_bench_reference.json` |")
    t.append(row(f"3: 256^3, DTU camera, 8192 rays, 1 GPU (M2/ray {c3['config']['samples_per_ray']['M2']:.1f})" if c3 else "3", c3, "r02_bench_cfg3.json"))
    t.append(row("2 at 2 GPUs (8192 rays per rank)", n2, "r02_bench_cfg2_n2.json",
                 extra=f": {n2['value'] / c2['value']:.2f}x of 1 GPU (both lines: final build)" if n2 else ""))
    t.append(row("2 at 8 GPUs", n8, "r02_bench_cfg2_n8.json", extra=f": {n8['value'] / c2s['value']:.2f}x of the 1-GPU line of the same build and visit ({c2s['ms_per_step']:.3f} ms, `r02_bench_cfg2_prev.json`)" if n8 else ""))
    t.append(row("4: 256^3, 8192 rays per rank, 8 GPUs", c4, "r02_bench_cfg4_n8.json"))
    t.append(row("5: 384^3, 800 x 800 image (640 K rays, 79 chunks), 1 GPU, inference", c5, "r02_bench_cfg5_n1.json", per=" / image"))
    t.append(row("5 at 8 GPUs (chunks dealt round-robin, one all-gather per output)", c58, "r02_bench_cfg5_n8.json", per=" / image",
                 extra=(f": {c58['value'] / c5p['value']:.2f}x of the 1-GPU line of the same build and visit ({M(c5p['value'])}, `r02_bench_cfg5_n1_prev.json`)"
                        if c58 and c5p else "")))
    t.append(row(f"coarse stage (8f row 1): 116^3, 8192 rays, 1 GPU (M2/ray {co['config']['samples_per_ray']['M2']:.1f})" if co else "coarse", co, "r02_bench_coarse.json"))
    ge = line("r02_bench_geometry.json")
    t.append(row(f"geometry_searching stage (8f row 1): 120^3, 6-ch k0, 8192 rays, 1 GPU (no MaskCache: M2/ray {ge['config']['samples_per_ray']['M2']:.1f})" if ge else "geometry_searching", ge, "r02_bench_geometry.json"))
    cfg_table = "\n".join(t)
    k = ["| kernel | ms / step | share | bound | achieved | fraction of the measured peak | DRAM bytes / launch (ncu) |", "|---|---|---|---|---|---|---|"]
    for r in c2["kernels"]:
        tr = r.get("ncu_dram_bytes_per_launch")
        tb = tr.get("bytes") if isinstance(tr, dict) else tr
        k.append(f"| `{r['name']}` | {r['ms_per_step']:.4f} | {100 * r['share']:.1f} % | {r.get('bound', '--')} | "
                 f"{str(r.get('achieved', '--')) + ' ' + r.get('unit', '') if r.get('achieved') is not None else '--'} | "
                 f"{r.get('frac', '--')} | {f'{tb / 1e6:.1f} MB' if tb else '--'} |")
    ktable = "\n".join(k)
    summary = (f"Headline: **{M(c2['value'])} rays/s** resident ({c2['ms_per_step']:.3f} ms per 8192-ray step), "
               f"**{M(c2['e2e']['value'])} rays/s** end to end"
               + (f" (host-fed loop with the next batch's upload and the previous step's loss read overlapped; "
                  f"{M(c2['e2e']['serial_value'])} with nothing overlapped)" if c2['e2e'].get('serial_value') else "")
               + (f"; the reference's own arithmetic on the same GPU runs at {M(go['value'])} rays/s ({c2['value'] / go['value']:.1f}x slower)" if go.get("value") else "")
               + (f"; 8 GPUs: {M(n8['value'])} rays/s = {n8['value'] / c2s['value']:.2f}x of the single-GPU line of the same build" if n8 else "") + ".")
    block = f"{summary}\n\n{cfg_table}\n\nPer-kernel roofline of the headline step:\n\n{ktable}\n"
    for fn in ("DESIGN.md", "README.md"):
        p = os.path.join(ROOT, fn)
        s = open(p).read()
        body = block if fn == "DESIGN.md" else f"{summary}\n\n{cfg_table}\n"
        s2 = re.sub(r"<!-- results:begin -->.*?<!-- results:end -->", "<!-- results:begin -->\n" + body.replace("\\", "\\\\") + "<!-- results:end -->", s, flags=re.S)
        open(p, "w").write(s2)
        print(fn, "updated" if s2 != s else "unchanged (markers missing?)")


if __name__ == "__main__":
    main()
