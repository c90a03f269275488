#!/usr/bin/env python
"""Golden vectors for the native ops, made by running THE REFERENCE'S OWN KERNELS
(oracle/_ref/*.so = model/cuda/{render_utils,adam_upd,total_variation} compiled for sm_100a by
oracle/build_ref.py) on a B200 over the seeded inputs of tests/native_cases.py.

Run on the GPU box:   gpurun -- 'python tests/golden/make_native_golden.py'
It writes gpurun_out/native_golden.npz; copy that file to tests/golden/native_golden.npz.
The CPU oracle (oracle/fgs_oracle.c) is pinned against it by tests/test_oracle_native.py.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import native_cases as nc  # noqa: E402
from oracle import ref_ext  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    ru = ref_ext.load("render_utils_cuda")
    ad = ref_ext.load("adam_upd_cuda")
    tv = ref_ext.load("total_variation_cuda")
    out = {}

    c = nc.sampling_case()
    r = ru.sample_pts_on_rays(cu(c["rays_o"]), cu(c["rays_d"]), cu(c["xyz_min"]), cu(c["xyz_max"]),
                              float(c["near"]), float(c["far"]), float(c["stepdist"]))
    for k, t in zip(["ray_pts", "mask_outbbox", "ray_id", "step_id", "N_steps", "t_min", "t_max"], r):
        out["samp_" + k] = t.cpu().numpy()
    s, d = ru.infer_ray_start_dir(cu(c["rays_o"]), cu(c["rays_d"]), r[5])
    out["samp_start"], out["samp_dir"] = s.cpu().numpy(), d.cpu().numpy()
    p, m = ru.sample_ndc_pts_on_rays(cu(c["rays_o"][:64]), cu(c["rays_d"][:64]), cu(c["xyz_min"]),
                                     cu(c["xyz_max"]), 17)
    out["ndc_pts"], out["ndc_mask"] = p.cpu().numpy(), m.cpu().numpy()
    out["bg_pts"] = ru.sample_bg_pts_on_rays(cu(c["rays_o"][:64]), cu(c["rays_d"][:64]), r[6][:64].contiguous(),
                                             0.5, 16).cpu().numpy()

    c = nc.alpha_case()
    w, T, last, i_s, i_e = ru.alpha2weight(cu(c["alpha"]), cu(c["ray_id"]), c["n_rays"])
    g = ru.alpha2weight_backward(cu(c["alpha"]), w, T, last, i_s, i_e, c["n_rays"], cu(c["grad_weights"]),
                                 cu(c["grad_last"]))
    for k, t in zip(["weight", "T", "last", "i_start", "i_end", "grad"], [w, T, last, i_s, i_e, g]):
        out["a2w_" + k] = t.cpu().numpy()

    c = nc.maskcache_case()
    out["mc_out"] = ru.maskcache_lookup(cu(c["world"]), cu(c["xyz"]), cu(c["scale"]), cu(c["shift"])).cpu().numpy()

    c = nc.raw2alpha_case()
    e, a = ru.raw2alpha(cu(c["density"]), float(c["shift"]), float(c["interval"]))
    out["r2a_exp"], out["r2a_alpha"] = e.cpu().numpy(), a.cpu().numpy()
    out["r2a_grad"] = ru.raw2alpha_backward(e, cu(c["grad_back"]), float(c["interval"])).cpu().numpy()
    e2, a2 = ru.raw2alpha_nonuni(cu(c["density"]), float(c["shift"]), cu(c["interval_pts"]))
    out["r2a_nu_alpha"] = a2.cpu().numpy()
    out["r2a_nu_grad"] = ru.raw2alpha_nonuni_backward(e2, cu(c["grad_back"]), cu(c["interval_pts"])).cpu().numpy()

    c = nc.adam_case()
    for mode, fn in (("dense", ad.adam_upd), ("masked", ad.masked_adam_upd), ("perlr", None)):
        p, m, v = cu(c["param"].copy()), cu(c["exp_avg"].copy()), cu(c["exp_avg_sq"].copy())
        for s, gr in enumerate(c["grads"]):
            if mode == "perlr":
                ad.adam_upd_with_perlr(p, cu(gr), m, v, cu(c["perlr"]), c["step0"] + s, c["beta1"], c["beta2"],
                                       c["lr"], c["eps"])
            else:
                fn(p, cu(gr), m, v, c["step0"] + s, c["beta1"], c["beta2"], c["lr"], c["eps"])
        out[f"adam_{mode}_p"], out[f"adam_{mode}_m"], out[f"adam_{mode}_v"] = (
            p.cpu().numpy(), m.cpu().numpy(), v.cpu().numpy())

    c = nc.tv_case()
    for dense in (True, False):
        g = cu(c["grad"].copy())
        tv.total_variation_add_grad(cu(c["param"]), g, c["wx"], c["wy"], c["wz"], dense)
        out[f"tv_{int(dense)}"] = g.cpu().numpy()
        g = cu(c["grad"].copy())
        tv.total_variation_add_grad_new(cu(c["param"]), g, cu(c["mask"]), c["wx"], c["wy"], c["wz"], dense)
        out[f"tvm_{int(dense)}"] = g.cpu().numpy()

    torch.cuda.synchronize()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    path = os.path.join(ROOT, "gpurun_out", "native_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
