#!/usr/bin/env python
"""Layout of the reference's fine-stage checkpoint (model/nerf_training.py:522-531), taken from THE REFERENCE'S OWN
objects instantiated in this container (same stubs as make_fine_golden.py): the keys / shapes / dtypes of
`model.state_dict()` after `model.set_sdf_mask()`, of `model.get_kwargs()` and `model.get_MaskCache_kwargs()`, and of
`optimizer.state_dict()` of `create_optimizer_or_freeze_model` after one step.

    python tests/golden/make_ckpt_golden.py      # needs /root/reference; writes ckpt_layout.json
"""
import json
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
REF = "/root/reference"


def main():
    import make_fine_golden as mk
    mk.install_stubs()
    sys.path.insert(0, REF)
    for n in ("tqdm",):
        try:
            __import__(n)
        except Exception:
            m = types.ModuleType(n)
            m.trange = range
            sys.modules[n] = m
    from fgs_nerf_b200 import scene
    G = 24
    sc = scene.make_scene(G=G, mask_G=20)
    mlin = torch.linspace(-1.2, 1.2, 20)
    mx, my, mz = torch.meshgrid(mlin, mlin, mlin, indexing="ij")
    sc.pre_pool_mask = ((((mx * mx + my * my + mz * mz).sqrt() - 0.6).abs() < 0.15).float() * 1e-3)[None, None]
    with tempfile.TemporaryDirectory() as td:
        model, ref_nerf = mk.build_reference_model(sc, td)
    # abs(sdf < 0.5) on a bool tensor raises on the CPU in torch 2.11 (SURVEY 8c quirk list); the CUDA result is the
    # mask itself as float
    try:
        model.set_sdf_mask()
    except Exception as e:
        print("set_sdf_mask on CPU:", type(e).__name__, e)
        import inspect
        print(inspect.getsource(model.set_sdf_mask))
    sd = model.state_dict()
    layout = {"model_state_dict": [[k, list(v.shape), str(v.dtype)] for k, v in sd.items()]}
    kw = model.get_kwargs()
    layout["model_kwargs"] = {k: (type(v).__name__ + (str(list(np.shape(v))) if hasattr(v, "shape") else "")) for k, v in kw.items()}
    mk_kw = model.get_MaskCache_kwargs()
    layout["MaskCache_kwargs"] = {k: (type(v).__name__ + (str(list(np.shape(v))) if hasattr(v, "shape") else "")) for k, v in mk_kw.items()}
    # optimizer: create_optimizer_or_freeze_model (nerf_training.py:9-37) executed from the reference's source
    from model import nerf_training as nt

    class Cfg(dict):
        __getattr__ = dict.__getitem__

    t = Cfg(scene.FINE_TRAIN)
    t["lrate_decay"] = 20
    t["skip_zero_grad_fields"] = ["density", "k0", "k1"]

    class L:
        def info(self, *a):
            pass

    opt = nt.create_optimizer_or_freeze_model(model, L(), t, global_step=0)
    for p in model.parameters():
        if p.requires_grad:
            p.grad = torch.zeros_like(p)
    opt.step()
    osd = opt.state_dict()
    layout["optimizer_param_groups"] = [{k: (v if not isinstance(v, (list, tuple)) or k == "betas" else len(v)) for k, v in g.items()}
                                        for g in osd["param_groups"]]
    st0 = osd["state"][0]
    layout["optimizer_state_entry"] = {k: (list(v.shape) if torch.is_tensor(v) else type(v).__name__) for k, v in st0.items()}
    layout["optimizer_state_count"] = len(osd["state"])
    json.dump(layout, open(os.path.join(HERE, "ckpt_layout.json"), "w"), indent=1)
    print(json.dumps(layout, indent=1)[:6000])


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("needs /root/reference")
    main()
