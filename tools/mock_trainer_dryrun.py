"""Dry run of the Python host paths (engine.scale_volume_grid, FineTrainer) WITHOUT a GPU: torch.cuda's
stream / event objects are replaced by dummies, the engine's device check is lifted and every C-ABI return code is
ignored (the launches fail with 'no CUDA device' and do nothing).  It only proves that the Python control flow,
attribute names, shapes and schedules hold together -- numerics are the GPU tests' job
(tests/test_training_gpu.py).  Development aid for a container without a GPU; not part of the product or the tests.

    python tools/mock_trainer_dryrun.py
"""
import contextlib
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class _Event:
    def record(self, *a):
        pass

    def synchronize(self):
        pass


class _Stream:
    cuda_stream = 0

    def wait_event(self, e):
        pass


torch.cuda.Event = _Event
torch.cuda.Stream = lambda **k: _Stream()
torch.cuda.current_stream = lambda *a: _Stream()
torch.cuda.device = lambda d: contextlib.nullcontext()
torch.cuda.stream = lambda s: contextlib.nullcontext()
torch.cuda.current_device = lambda: 0
torch.Tensor.pin_memory = lambda self: self

import fgs_nerf_b200  # noqa: E402,F401
from fgs_nerf_b200 import _lib  # noqa: E402

_lib.check_input = lambda *a, **k: None
_lib.stream = lambda: 0
_lib.check = lambda rc: None
src = open(os.path.join(ROOT, "fgs_nerf_b200", "engine.py")).read().replace('if self.device.type != "cuda":', "if False:").replace("assert grid.is_cuda and", "assert")
mod = types.ModuleType("fgs_nerf_b200.engine")
mod.__package__ = "fgs_nerf_b200"
mod.__file__ = os.path.join(ROOT, "fgs_nerf_b200", "engine.py")
sys.modules["fgs_nerf_b200.engine"] = mod
exec(compile(src, mod.__file__, "exec"), mod.__dict__)
fgs_nerf_b200.engine = mod

from fgs_nerf_b200 import scene, training  # noqa: E402

engine = mod
dev = "cpu"
N, n_total = 64, 500
sc = scene.make_scene(G=24, mask_G=20)
rays = scene.make_rays(N=n_total, H=50, W=50, n_views=4)

out = engine.grid_resize(sc.sdf, (30, 30, 30))
assert tuple(out.shape) == (1, 1, 30, 30, 30)
k0cl = sc.k0.contiguous(memory_format=torch.channels_last_3d)
out = engine.grid_resize(k0cl, (30, 31, 32))
assert tuple(out.shape) == (1, 12, 30, 31, 32) and not out.is_contiguous()
o = engine.gather_rays(*rays, torch.arange(10))
assert o[0].shape == (10, 3)

w = engine.gaussian_3dconv_weights(5, 1)
assert tuple(w.shape) == (5, 5, 5) and abs(float(w.sum()) - 1) < 1e-6
assert engine.smooth_conv(sc.sdf, 5, 1).shape == sc.sdf.shape
e0 = engine.FineStepEngine(sc, N, dev)
e0.set_nonempty_mask()
p0 = e0.sdf.data_ptr()
e0.init_sdf_from_sdf(torch.randn(1, 1, 10, 10, 10), smooth=False, reduce=0.3)
e0.init_sdf_from_sdf(sc.sdf, smooth=True, reduce=0.5, ksize=3, sigma=1.0, smooth_scale=False)
assert e0.sdf.data_ptr() == p0
u = e0.extract_fields(torch.tensor([-1.0] * 3), torch.tensor([1.0] * 3), 20)
assert tuple(u.shape) == (20, 20, 20)
from fgs_nerf_b200 import extract_geometry  # noqa: E402
v = extract_geometry.extract_fields(torch.tensor([-1.0] * 3), torch.tensor([1.0] * 3), 20, lambda p: p[:, 0], N=8)
assert v.shape == (20, 20, 20) and abs(v[-1, 0, 0] - 1.0) < 1e-6

ratio = (32 / 24) ** 3
cfg = dict(scene.FINE_TRAIN, tv_every=1, N_rand=N, lrate_decay=20, pg_scale=[5002], scale_ratio=ratio,
           decay_step_module={5000: dict(sdf=0.1)}, tv_updates={5001: dict(sdf_tv=0.3)}, s_updates={5001: dict(s_ratio=40)})
eng = engine.FineStepEngine(sc, N, dev, train_cfg=dict(scene.FINE_TRAIN))
eng.set_nonempty_mask()
tr = training.FineTrainer(eng, *rays, cfg_train=cfg, start=5000, rng=np.random.RandomState(1))
tr.run(4, i_print=2, log=lambda s: print("  log:", s))
st = tr.stats()
print("steps", st["steps"], "lr", st["lr"])
print("engine grid", tr.engine.X, tr.engine.Y, tr.engine.Z, "adam_step", tr.engine.adam_step, "num_voxels", tr.engine.num_voxels)
assert tr.engine.train_cfg["sdf_tv"] == 0.3 and tr.engine.cfg["s_ratio"] == 40
assert tr.engine.adam_step == 3
import tempfile  # noqa: E402

from fgs_nerf_b200 import checkpoint  # noqa: E402

ck = checkpoint.checkpoint_dict(tr.engine, tr.global_step)
assert len(ck["optimizer_state_dict"]["state"]) == 18 and len(ck["model_kwargs"]) == 22
with tempfile.TemporaryDirectory() as td:
    checkpoint.save_checkpoint(tr.engine, tr.global_step, os.path.join(td, "fine_last.tar"))
    e2, start = checkpoint.engine_from_checkpoint(os.path.join(td, "fine_last.tar"), N, dev, train_cfg=dict(tr.engine.train_cfg))
assert start == tr.global_step and e2.adam_step == tr.engine.adam_step and e2.lr == tr.engine.lr
assert torch.equal(e2.k0, tr.engine.k0) and torch.equal(e2.weights[4], tr.engine.weights[4])
assert torch.equal(e2.nonempty_mask, tr.engine.nonempty_mask)
from fgs_nerf_b200 import nerf as nerf_mod  # noqa: E402

mc = nerf_mod.MaskCache(mask_cache_thres=1e-3, sdf_mask=sc.mask_grid, xyz_min=sc.mask_xyz_min, xyz_max=sc.mask_xyz_max)
kw = dict(xyz_min=[-1, -1, -1], xyz_max=[1, 1, 1], num_voxels=24 ** 3, num_voxels_base=24 ** 3, mask_cache=mc,
          fast_color_thres=1e-4, ref=True, grad_feat=(0.5, 1.0, 1.5, 2.0), sdf_feat=(0.5, 1.0, 1.5, 2.0), s_ratio=50,
          s_start=0.05, stage="fine", n_rays=N, device=dev)
model = nerf_mod.nerf(**kw)
assert tuple(model.sdf.grid.shape) == (1, 1, 24, 24, 24) and tuple(model.k0.grid.shape) == (1, 12, 24, 24, 24)
rk = dict(near=2.0, far=6.0, stepsize=0.5, bg=1.0, render_depth=True, render_grad=True)
full = model(rays[0][:N], rays[1][:N], rays[2][:N], **rk)
part = model(rays[0][:10], rays[1][:10], rays[2][:10], **rk)
ref_keys = ["alphainv_cum", "weights", "ray_id", "viewdirs", "rgb_marched", "sigmoid_rgb", "normal_marched", "normal",
            "raw_alpha", "raw_rgb", "depth", "disp", "mask", "mask_outbbox", "gradient", "s_val"]
print("missing keys:", [k for k in ref_keys if k not in full], [k for k in ref_keys if k not in part])
assert part["rgb_marched"].shape == (10, 3) and part["alphainv_cum"].shape == (10,)
sd = model.state_dict()
model.load_state_dict(sd)
model.scale_volume_grid(model.num_voxels * (32 / 24) ** 3)
assert model.world_size.tolist() == [32, 32, 32]
try:
    nerf_mod.nerf(**dict(kw, stage="coarse"))
    raise SystemExit("coarse stage must be refused")
except NotImplementedError:
    pass
print("dry run ok")
