"""The reference's checkpoint format for the fine stage (model/nerf_training.py:522-531, read back by
model/utils.py:26-56): a dict with `global_step`, `model_kwargs`, `MaskCache_kwargs`, `model_state_dict`,
`optimizer_state_dict`, written with torch.save.  Key names, order, shapes and dtypes follow the reference's own
objects (tests/golden/ckpt_layout.json, dumped from them by tests/golden/make_ckpt_golden.py), so a checkpoint written
here loads into the reference's `nerf` (`load_state_dict(..., strict=False)`, as utils.py:47 calls it) and one written
by the reference loads here.

Not written: `grad_conv.*` / `tv_smooth_conv.*` -- fixed stencils the reference's constructor re-creates
(model/nerf.py:225-257), frozen, never trained; the reference loads with strict=False.
k0 is stored in the reference's contiguous [1,12,X,Y,Z] layout (the engine keeps it channels-last in memory).
"""
from collections import OrderedDict

import torch

from .engine import K0_DIM, FineStepEngine
from .scene import Scene, mlp_dims

# nn.Sequential indices of the Linear layers (model/nerf.py:125-142: Linear+ReLU, [Sequential(Linear, ReLU)] x2, Linear)
_LIN = ("0", "2.0", "3.0", "4")
_LIN3 = ("0", "2.0", "3")                            # the coarse stage's 3-layer refnet
_GROUPS = ("k0", "sdf", "rgbnet", "refnet")          # order of create_optimizer_or_freeze_model's param groups


def _mlp_items(eng):
    """(key, weight view [out,in], bias, engine layer index) in the reference's state_dict order: refnet, then rgbnet."""
    out = []
    if hasattr(eng, "biases"):                       # coarse.CoarseStepEngine: refnet only
        return [(f"refnet.{idx}", j) for j, idx in enumerate(_LIN3)]
    for name, base in (("refnet", 4), ("rgbnet", 0)):
        for j, idx in enumerate(_LIN):
            out.append((f"{name}.{idx}", base + j))
    return out


def model_state_dict(eng, global_step=None):
    c = eng.cfg
    dev = eng.device
    sd = OrderedDict()
    sv = eng.s_val(global_step) if global_step is not None else 0.0
    sd["s_val"] = torch.full((1,), float(sv), dtype=torch.float32, device=dev)
    sd["xyz_min"] = eng.xyz_min.to(dev)
    sd["xyz_max"] = eng.xyz_max.to(dev)
    for key, n in (("posfreq", c["posbase_pe"]), ("viewfreq", c["viewbase_pe"]), ("reffreq", c["refbase_pe"])):
        sd[key] = torch.FloatTensor([(2 ** i) for i in range(n)]).to(dev)           # model/nerf.py:98-104
    if eng.nonempty_mask is not None:
        sd["nonempty_mask"] = eng.nonempty_mask.reshape(eng.sdf.shape).clone()
    sd["sdf.grid"] = eng.sdf.detach().clone()
    sd["k0.grid"] = eng.k0.detach().contiguous().clone()                            # NCDHW, as the reference stores it
    for key, i in _mlp_items(eng):
        sd[key + ".weight"] = eng.weights[i].detach().clone().contiguous()
        sd[key + ".bias"] = (eng.biases[i] if hasattr(eng, "biases") else eng.Bt[i]).detach().clone()
    if eng.mask_grid is not None:                    # (the geometry_searching stage has no MaskCache)
        sd["mask_cache.xyz_min"] = eng.mask_xyz_min.to(dev)
        sd["mask_cache.xyz_max"] = eng.mask_xyz_max.to(dev)
        sd["mask_cache.sdf_mask"] = eng.mask_grid.clone()
    # nerf.set_sdf_mask (model/nerf.py:181-186), what the NEXT stage's MaskCache reads (nerf.py:1199)
    sd["sdf_mask.grid"] = (eng.sdf.detach() < 0.5).float() * 1e-3
    return sd


def optimizer_state_dict(eng):
    """torch.optim.Optimizer.state_dict() of the reference's MaskedAdam (model/adam.py:167-221): params numbered in
    group order k0, sdf, rgbnet (weight, bias per layer), refnet."""
    eng._ensure_adam()
    params, groups, k = [], [], 0
    coarse = hasattr(eng, "biases")
    per_group = {"k0": [("k0", None)], "sdf": [("sdf", None)],
                 "rgbnet": [(kind, i) for i in range(0, 4) for kind in ("W", "B")],
                 "refnet": [(kind, i) for i in (range(3) if coarse else range(4, 8)) for kind in ("W", "B")]}
    skip = tuple(eng.train_cfg.get("skip_zero_grad_fields", ("k0",))) if coarse else ("k0",)
    for name in (("k0", "sdf", "refnet") if coarse else _GROUPS):
        ids = []
        for kind, i in per_group[name]:
            if i is None:
                m, v = eng.adam[kind]
                if kind == "k0" and m.shape[1] != eng.k0.shape[1]:       # (coarse stages: 12-channel storage of a narrower grid)
                    m, v = m[:, :eng.k0.shape[1]], v[:, :eng.k0.shape[1]]
                m, v = m.contiguous().clone(), v.contiguous().clone()
            elif kind == "W":       # (the engines store the layers zero-padded: cut back to nn.Linear's shape)
                o, n_in = eng.weights[i].shape
                m, v = (t[:o, :n_in].contiguous().clone() for t in eng.adam["W"][i])
            else:
                o = eng.weights[i].shape[0]
                m, v = (t[:o].clone() for t in eng.adam["B"][i])
            params.append({"step": int(eng.adam_step), "exp_avg": m, "exp_avg_sq": v})
            ids.append(k)
            k += 1
        groups.append({"lr": eng.lr[name], "name": name, "skip_zero_grad": name in skip, "betas": (0.9, 0.99), "eps": 1e-8,
                       "params": ids})
    state = {i: p for i, p in enumerate(params)} if eng.adam_step > 0 else {}
    return {"state": state, "param_groups": groups}


def _coarse_model_kwargs(eng):
    """nerf.get_kwargs (model/nerf.py:309-329) of a stage='coarse' model: no grad / sdf features, rgbnet_dim is still
    computed by the constructor (:108-112) although no rgbnet is built."""
    c = eng.cfg
    rgb_dim = (3 + 3 * c["posbase_pe"] * 2) + c["k0_dim"] + 3 + 1 + (3 + 3 * c["viewbase_pe"] * 2)
    return {
        "xyz_min": eng.xyz_min.cpu().numpy(), "xyz_max": eng.xyz_max.cpu().numpy(),
        "num_voxels": eng.num_voxels, "num_voxels_base": eng.num_voxels_base, "voxel_size": torch.tensor(float(eng.voxel_size)),
        "nearest": False, "k0_dim": c["k0_dim"], "grad_feat": (), "sdf_feat": (), "center_sdf": True,
        "fast_color_thres": float(c["fast_color_thres"]), "stage": eng.stage, "ref": True, "use_viewdir": True,
        "s_ratio": c["s_ratio"], "s_start": c["s_start"],
        "rgbnet_dim": rgb_dim, "rgbnet_width": c.get("rgbnet_width", 192), "rgbnet_depth": c.get("rgbnet_depth", 3),
        "refnet_dim": eng.in_dim, "refnet_width": c["refnet_width"], "refnet_depth": c["refnet_depth"],
    }


def checkpoint_dict(eng, global_step):
    c = eng.cfg
    if hasattr(eng, "biases"):                      # coarse.CoarseStepEngine
        ext = (eng.xyz_max - eng.xyz_min).float()
        voxel_size_base = (ext.prod() / eng.num_voxels_base).pow(1 / 3)
        mc = {"xyz_min": eng.xyz_min.cpu().numpy(), "xyz_max": eng.xyz_max.cpu().numpy(),
              "voxel_size_ratio": torch.tensor(float(eng.voxel_size)) / voxel_size_base, "nearest": False}
        return {"global_step": int(global_step), "model_kwargs": _coarse_model_kwargs(eng), "MaskCache_kwargs": mc,
                "model_state_dict": model_state_dict(eng, global_step), "optimizer_state_dict": optimizer_state_dict(eng)}
    rgb_dim, ref_dim = mlp_dims(c["k0_dim"], c["posbase_pe"], c["viewbase_pe"], c["refbase_pe"], len(c["grad_feat"]),
                                c["rgbnet_width"])
    kwargs = {   # nerf.get_kwargs (model/nerf.py:309-329)
        "xyz_min": eng.xyz_min.cpu().numpy(), "xyz_max": eng.xyz_max.cpu().numpy(),
        "num_voxels": eng.num_voxels, "num_voxels_base": eng.num_voxels_base, "voxel_size": torch.tensor(float(eng.voxel_size)),
        "nearest": False, "k0_dim": c["k0_dim"], "grad_feat": tuple(c["grad_feat"]), "sdf_feat": tuple(c["sdf_feat"]),
        "center_sdf": True, "fast_color_thres": float(c["fast_color_thres"]), "stage": "fine", "ref": True,
        "use_viewdir": True, "s_ratio": c["s_ratio"], "s_start": c["s_start"],
        "rgbnet_dim": rgb_dim, "rgbnet_width": c["rgbnet_width"], "rgbnet_depth": c["rgbnet_depth"],
        "refnet_dim": ref_dim, "refnet_width": c["refnet_width"], "refnet_depth": c["refnet_depth"],
    }
    # voxel_size_ratio = voxel_size / voxel_size_base, voxel_size_base from num_voxels_base (model/nerf.py:68, 303)
    ext = (eng.xyz_max - eng.xyz_min).float()
    voxel_size_base = (ext.prod() / eng.num_voxels_base).pow(1 / 3)
    mc = {"xyz_min": eng.xyz_min.cpu().numpy(), "xyz_max": eng.xyz_max.cpu().numpy(),      # get_MaskCache_kwargs (:331-337)
          "voxel_size_ratio": torch.tensor(float(eng.voxel_size)) / voxel_size_base, "nearest": False}
    return {"global_step": int(global_step), "model_kwargs": kwargs, "MaskCache_kwargs": mc,
            "model_state_dict": model_state_dict(eng, global_step), "optimizer_state_dict": optimizer_state_dict(eng)}


def save_checkpoint(eng, global_step, path):
    torch.save(checkpoint_dict(eng, global_step), path)


def engine_from_checkpoint(ckpt, n_rays, device="cuda", model_cfg=None, train_cfg=None, no_reload_optimizer=False, **kw):
    """utils.load_checkpoint (model/utils.py:42-56) for the fine stage: a FineStepEngine holding the checkpoint's
    parameters (and, unless no_reload_optimizer, its MaskedAdam state and learning rates).  `ckpt`: path or dict.
    Returns (engine, start)."""
    from .scene import FINE_MODEL
    if not isinstance(ckpt, dict):
        ckpt = torch.load(ckpt, map_location="cpu", weights_only=False)
    sd = ckpt["model_state_dict"]
    cfg = dict(model_cfg if model_cfg is not None else FINE_MODEL)
    for k in ("k0_dim", "s_ratio", "s_start", "fast_color_thres", "rgbnet_width", "rgbnet_depth", "refnet_width",
              "refnet_depth"):
        if k in ckpt.get("model_kwargs", {}):
            cfg[k] = ckpt["model_kwargs"][k]
    sdf, k0 = sd["sdf.grid"].float(), sd["k0.grid"].float()
    if k0.shape[1] != K0_DIM or tuple(k0.shape[2:]) != tuple(sdf.shape[2:]):
        raise RuntimeError(f"checkpoint grids {tuple(sdf.shape)} / {tuple(k0.shape)} are not a fine-stage pair")
    xyz_min, xyz_max = sd["xyz_min"].float().cpu(), sd["xyz_max"].float().cpu()
    mk = ckpt.get("model_kwargs", {})
    if "voxel_size" in mk:
        voxel_size = float(mk["voxel_size"])
    else:
        voxel_size = float(((xyz_max - xyz_min).prod() / sdf[0, 0].numel()).pow(1 / 3))
    lin = lambda name: [(sd[f"{name}.{i}.weight"].float(), sd[f"{name}.{i}.bias"].float()) for i in _LIN]  # noqa: E731
    sc = Scene(G=int(sdf.shape[2]), xyz_min=xyz_min, xyz_max=xyz_max, voxel_size=voxel_size, sdf=sdf, k0=k0,
               mask_grid=sd["mask_cache.sdf_mask"].float(), mask_xyz_min=sd["mask_cache.xyz_min"].float().cpu(),
               mask_xyz_max=sd["mask_cache.xyz_max"].float().cpu(), mask_thres=cfg["mask_cache_thres"],
               rgbnet=lin("rgbnet"), refnet=lin("refnet"), cfg=cfg)
    eng = FineStepEngine(sc, n_rays, device, train_cfg=train_cfg, **kw)
    if "num_voxels" in mk:
        eng.num_voxels = mk["num_voxels"]
    eng.num_voxels_base = mk.get("num_voxels_base", eng.num_voxels)
    if "nonempty_mask" in sd:
        eng.nonempty_mask = sd["nonempty_mask"].to(eng.device).reshape(eng.X, eng.Y, eng.Z).contiguous()
        eng.n_nonempty = int(eng.nonempty_mask.sum().item())
    start = int(ckpt.get("global_step", 0))
    osd = ckpt.get("optimizer_state_dict")
    if not no_reload_optimizer and osd and osd.get("state"):
        load_optimizer_state_dict(eng, osd)
    return eng, start


def load_optimizer_state_dict(eng, osd):
    eng._ensure_adam()
    st = osd["state"]
    ids = {g["name"]: list(g["params"]) for g in osd["param_groups"]}
    for g in osd["param_groups"]:
        if g["name"] in eng.lr:
            eng.lr[g["name"]] = float(g["lr"])
    steps = {int(st[i]["step"]) for i in st}
    if len(steps) != 1:
        raise RuntimeError(f"per-parameter Adam steps differ ({sorted(steps)}): the fused optimiser keeps one step count")
    eng.adam_step = steps.pop()

    def put(dst, src):
        dst.zero_()
        view = dst[:, :src.shape[1]] if dst.dim() == 2 and dst.shape != src.shape else dst
        view.copy_(src.to(dst.device))

    for name, key in (("k0", "k0"), ("sdf", "sdf")):
        e = st[ids[name][0]]
        put(eng.adam[key][0], e["exp_avg"])
        put(eng.adam[key][1], e["exp_avg_sq"])
    for name, base in (("rgbnet", 0), ("refnet", 4)):
        for j in range(4):
            w, b = st[ids[name][2 * j]], st[ids[name][2 * j + 1]]
            put(eng.adam["W"][base + j][0], w["exp_avg"])
            put(eng.adam["W"][base + j][1], w["exp_avg_sq"])
            put(eng.adam["B"][base + j][0], b["exp_avg"])
            put(eng.adam["B"][base + j][1], b["exp_avg_sq"])



def coarse_engine_from_checkpoint(ckpt, n_rays, device="cuda", model_cfg=None, train_cfg=None, no_reload_optimizer=False, **kw):
    """utils.load_checkpoint (model/utils.py:42-56) for a stage='coarse' checkpoint (also one the reference wrote):
    a coarse.CoarseStepEngine holding its parameters and, unless no_reload_optimizer, its MaskedAdam state.
    Returns (engine, start)."""
    from .coarse import COARSE_MODEL, CoarseStepEngine
    if not isinstance(ckpt, dict):
        ckpt = torch.load(ckpt, map_location="cpu", weights_only=False)
    sd, mk = ckpt["model_state_dict"], ckpt.get("model_kwargs", {})
    cfg = dict(model_cfg if model_cfg is not None else COARSE_MODEL)
    for k in ("k0_dim", "s_ratio", "s_start", "fast_color_thres", "refnet_width", "refnet_depth", "stage"):
        if k in mk:
            cfg[k] = mk[k]
    for key, name in (("posfreq", "posbase_pe"), ("viewfreq", "viewbase_pe"), ("reffreq", "refbase_pe")):
        if key in sd:
            cfg[name] = int(sd[key].numel())
    sdf, k0 = sd["sdf.grid"].float(), sd["k0.grid"].float()
    xyz_min, xyz_max = sd["xyz_min"].float().cpu(), sd["xyz_max"].float().cpu()
    voxel_size = float(mk["voxel_size"]) if "voxel_size" in mk else float(((xyz_max - xyz_min).prod() / sdf[0, 0].numel()).pow(1 / 3))
    refnet = [(sd[f"refnet.{i}.weight"].float(), sd[f"refnet.{i}.bias"].float()) for i in _LIN3]
    cfg["k0_dim"] = int(k0.shape[1])
    if "mask_cache.sdf_mask" in sd:
        mg, mmin, mmax = sd["mask_cache.sdf_mask"].float(), sd["mask_cache.xyz_min"].float().cpu(), sd["mask_cache.xyz_max"].float().cpu()
    else:                                            # geometry_searching: no MaskCache
        cfg["stage"] = "geometry_searching"
        mg, mmin, mmax = torch.zeros(1, 1, 2, 2, 2), xyz_min, xyz_max
    sc = Scene(G=int(sdf.shape[2]), xyz_min=xyz_min, xyz_max=xyz_max, voxel_size=voxel_size, sdf=sdf, k0=k0,
               mask_grid=mg, mask_xyz_min=mmin, mask_xyz_max=mmax, mask_thres=cfg["mask_cache_thres"], cfg=cfg)
    eng = CoarseStepEngine(sc, refnet, n_rays, device, cfg=cfg, train_cfg=train_cfg, **kw)
    if "num_voxels" in mk:
        eng.num_voxels = mk["num_voxels"]
    eng.num_voxels_base = mk.get("num_voxels_base", eng.num_voxels)
    if "nonempty_mask" in sd:
        eng.nonempty_mask = sd["nonempty_mask"].to(eng.device).reshape(eng.X, eng.Y, eng.Z).contiguous()
        eng.n_nonempty = int(eng.nonempty_mask.sum().item())
    osd = ckpt.get("optimizer_state_dict")
    if not no_reload_optimizer and osd and osd.get("state"):
        eng._ensure_adam()
        st = osd["state"]
        ids = {g["name"]: list(g["params"]) for g in osd["param_groups"]}
        for g in osd["param_groups"]:
            if g["name"] in eng.lr:
                eng.lr[g["name"]] = float(g["lr"])
        steps = {int(st[i]["step"]) for i in st}
        if len(steps) != 1:
            raise RuntimeError(f"per-parameter Adam steps differ ({sorted(steps)}): the fused optimiser keeps one step count")
        eng.adam_step = steps.pop()

        def put(dst, src):
            dst.zero_()
            view = dst[tuple(slice(0, n) for n in src.shape)]
            view.copy_(src.to(dst.device))

        for name in ("k0", "sdf"):
            e = st[ids[name][0]]
            put(eng.adam[name][0], e["exp_avg"])             # (k0: into the first k0_dim channels of the 12-channel storage)
            put(eng.adam[name][1], e["exp_avg_sq"])
        for j in range(3):
            w, b = st[ids["refnet"][2 * j]], st[ids["refnet"][2 * j + 1]]
            put(eng.adam["W"][j][0], w["exp_avg"]); put(eng.adam["W"][j][1], w["exp_avg_sq"])
            put(eng.adam["B"][j][0], b["exp_avg"]); put(eng.adam["B"][j][1], b["exp_avg_sq"])
    return eng, int(ckpt.get("global_step", 0))
