"""Digest of an `ncu --set full` report (read here, no GPU): per kernel the duration, DRAM bytes per launch, DRAM /
tensor-pipe / warp-occupancy percentages, registers -> profiles/<tag>_ncu_full.txt and profiles/<tag>_ncu_traffic.json
(the `traffic` figures bench.py reports).   usage: python tools/ncu_digest.py gpurun_out/r2_full.ncu-rep r02"""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "lts__t_sector_hit_rate.pct",
           "sm__inst_executed.sum", "sm__cycles_active.avg", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "launch__grid_size", "launch__block_size"]
# launch name in the library's own per-launch timing (fgs_prof_*) for each kernel function
LABEL = [("mlp_chain2_kernel", "mlp_chain_fwd"), ("mlp_bwd_planes_kernel", "mlp_chain_bwd"), ("mlp_dw2_kernel", "mlp_dw"),
         ("dw_small_kernel", "mlp_dw_small"), ("weight_split_kernel", "weight_split"),
         ("sample1_bwd_kernel", "fine level1_bwd"), ("sample1_kernel", "fine level1"), ("feat_bwd_kernel", "fine feat_bwd"),
         ("feat_kernel", "fine feat"), ("march_kernel", "fine march"), ("expand1_kernel", "fine expand1"),
         ("compact2_kernel", "fine compact2"), ("composite_kernel", "fine composite"),
         ("adam_bricks_kernel", "adam_bricks"), ("adam_dense_zero_kernel", "adam_dense_zero"), ("adam_multi", "adam_upd_multi"),
         ("brick_ulist_kernel", "brick_ulist"), ("weights1_bwd_kernel", "fine weights1_bwd"), ("weights1_kernel", "fine weights1"),
         ("tv_grad_grid", "smooth_tv grad grid"), ("tv_err", "smooth_tv err"), ("tv_scatter", "smooth_tv scatter"),
         ("tv_kernel", "total_variation_add_grad_v4")]


def unit_scale(unit):
    u = unit.strip().lower()
    return {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "usecond": 1.0,
            "nsecond": 1e-3, "msecond": 1e3, "second": 1e6}.get(u, 1.0)


def main(rep, tag):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    name_c = col.get("Kernel Name", col.get("Name"))
    agg = {}
    for r in data:
        if len(r) != len(hdr):
            continue
        k = r[name_c]
        d = agg.setdefault(k, {"n": 0})
        d["n"] += 1
        for m in METRICS:
            if m in col and r[col[m]] not in ("", "n/a"):
                v = float(r[col[m]].replace(",", "")) * (unit_scale(units[col[m]]) if ("bytes" in m or "duration" in m) else 1.0)
                d.setdefault(m, []).append(v)
    lines = [f"# ncu --set full --clock-control none of `bench.py --steps 3 --warmup 3` ({rep}); averages per launch; durations are "
             f"cold-cache, serialised (compare shares, not absolutes)"]
    traffic = {}
    for k, d in sorted(agg.items(), key=lambda kv: -sum(kv[1].get("gpu__time_duration.sum", [0]))):
        avg = lambda m: (sum(d[m]) / len(d[m])) if m in d else float("nan")
        rd, wr = avg("dram__bytes_read.sum"), avg("dram__bytes_write.sum")
        short = re.sub(r"\(.*", "", k)
        lines.append(f"{short}\n    launches {d['n']}  duration {avg('gpu__time_duration.sum'):.1f} us  DRAM read {rd / 1e6:.2f} MB  write {wr / 1e6:.2f} MB"
                     f"  DRAM thr {avg(METRICS[3]):.1f} %  tensor pipe {avg(METRICS[4]):.1f} % of active / {avg(METRICS[5]):.1f} % of elapsed"
                     f"  warps active {avg(METRICS[6]):.1f} %  issue active {avg(METRICS[11]):.1f} %  L2 hit {avg(METRICS[8]):.1f} %  regs {avg(METRICS[7]):.0f}"
                     f"  grid {avg(METRICS[12]):.0f} x {avg(METRICS[13]):.0f}")
        for pat, label in LABEL:
            if pat in k and label not in traffic:
                traffic[label] = {"bytes": rd + wr, "read": rd, "write": wr, "ncu_duration_us": avg("gpu__time_duration.sum"),
                                  "source": f"profiles/{tag}_ncu_full.txt ({short})"}
    open(os.path.join(ROOT, "profiles", f"{tag}_ncu_full.txt"), "w").write("\n".join(lines) + "\n")
    json.dump(traffic, open(os.path.join(ROOT, "profiles", f"{tag}_ncu_traffic.json"), "w"), indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
