"""TEST INFRASTRUCTURE (oracle): CPU restatements of the between-step operators of the training loop
(SURVEY.md 8f rows 2-3).  Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may import this.

  grid_resize_trilinear   DenseGrid.scale_volume_grid (model/grid.py:101-106) = F.interpolate(grid, size,
                          mode='trilinear', align_corners=True); the arithmetic restated here is ATen's
                          upsample_trilinear3d (PyTorch 2.11, not part of /root/reference): scale =
                          (float)(in-1)/(out-1), src = scale*dst, i0 = trunc(src), l1 = src - i0, l0 = 1 - l1,
                          value = d0*(h0*(w0*v000 + w1*v001) + h1*(...)) + d1*(...).
                          Pinned against torch's own F.interpolate (the reference's call) in
                          tests/test_training_host.py.
  gaussian_3dconv         nerf._gaussian_3dconv (model/nerf.py:260-272), pinned against the reference's own method
                          executed from its source (tests/golden/make_stage_golden.py); init_sdf_from_sdf (:280-296).
  gather_rays             the four `x_tr[sel_i]` selections of model/nerf_training.py:253-258.
  RefSchedule             python restatement of the loop's schedule block (nerf_training.py:244-254, 386-433),
                          pinned against tests/golden/schedule_golden.npz (the reference's block executed
                          from its source, tests/golden/make_schedule_golden.py).
"""
import math

import numpy as np


def _axis(n_in, n_out):
    scale = np.float32(n_in - 1) / np.float32(n_out - 1) if n_out > 1 else np.float32(0)
    r = (scale * np.arange(n_out, dtype=np.float32)).astype(np.float32)
    i0 = r.astype(np.int64)
    ip = (i0 < n_in - 1).astype(np.int64)
    l1 = (r - i0.astype(np.float32)).astype(np.float32)
    l0 = (np.float32(1) - l1).astype(np.float32)
    return i0, i0 + ip, l0, l1


def grid_resize_trilinear(src, size):
    """src [C,X,Y,Z] float32 -> [C,X2,Y2,Z2] float32."""
    src = np.asarray(src, dtype=np.float32)
    C, X, Y, Z = src.shape
    X2, Y2, Z2 = (int(v) for v in size)
    x0, x1, a0, a1 = _axis(X, X2)
    y0, y1, b0, b1 = _axis(Y, Y2)
    z0, z1, c0, c1 = _axis(Z, Z2)
    a0, a1 = a0[None, :, None, None], a1[None, :, None, None]
    b0, b1 = b0[None, None, :, None], b1[None, None, :, None]
    c0, c1 = c0[None, None, None, :], c1[None, None, None, :]

    def v(xi, yi, zi):
        return src[:, xi][:, :, yi][:, :, :, zi]

    f = np.float32
    lo = (b0 * (c0 * v(x0, y0, z0) + c1 * v(x0, y0, z1)).astype(f) + b1 * (c0 * v(x0, y1, z0) + c1 * v(x0, y1, z1)).astype(f)).astype(f)
    hi = (b0 * (c0 * v(x1, y0, z0) + c1 * v(x1, y0, z1)).astype(f) + b1 * (c0 * v(x1, y1, z0) + c1 * v(x1, y1, z1)).astype(f)).astype(f)
    return (a0 * lo + a1 * hi).astype(f)


def gaussian_3dconv_weights(ksize=3, sigma=1):
    """model/nerf.py:261-268: float64 exp -> fp32 -> divided by the fp32 sum."""
    r = np.arange(-(ksize // 2), ksize // 2 + 1, 1)
    xx, yy, zz = np.meshgrid(r, r, r)
    k = np.exp(-(xx ** 2 + yy ** 2 + zz ** 2) / (2 * sigma ** 2)).astype(np.float32)
    return (k / k.sum(dtype=np.float32)).astype(np.float32)


def gaussian_3dconv(vol, ksize=3, sigma=1, pre_div=1.0, post_div=1.0):
    """nerf._gaussian_3dconv(ksize, sigma) applied to vol [X,Y,Z] (model/nerf.py:260-272): Conv3d(1, 1, k,
    padding=k//2, padding_mode='replicate'), zero bias; cross-correlation like torch's conv.  fp32 accumulation in
    (a, b, c) order.  Pinned against the reference's own method executed from its source
    (tests/golden/make_stage_golden.py)."""
    f = np.float32
    w = gaussian_3dconv_weights(ksize, sigma)
    h = ksize // 2
    v = np.asarray(vol, dtype=f)
    if pre_div != 1.0:
        v = (v * (f(1) / f(pre_div))).astype(f)       # torch CUDA: tensor / python scalar = x * (1.0f / s)
    p = np.pad(v, h, mode="edge")
    X, Y, Z = v.shape
    acc = np.zeros_like(v)
    for a in range(ksize):
        for b in range(ksize):
            for c in range(ksize):
                acc = (acc + w[a, b, c] * p[a:a + X, b:b + Y, c:c + Z]).astype(f)
    if post_div != 1.0:
        acc = (acc * (f(1) / f(post_div))).astype(f)
    return acc


def init_sdf_from_sdf(sdf0, world_size, smooth=False, reduce=1.0, ksize=3, sigma=1.0, smooth_scale=True):
    """model/nerf.py:280-296 on a [X0,Y0,Z0] volume -> [X,Y,Z]."""
    v = np.asarray(sdf0, dtype=np.float32)
    if tuple(v.shape) != tuple(world_size):
        v = grid_resize_trilinear(v[None], world_size)[0]
    if smooth:
        v = gaussian_3dconv(v, ksize, sigma, pre_div=reduce, post_div=reduce)
    else:
        v = (v * (np.float32(1) / np.float32(reduce))).astype(np.float32)
    if smooth_scale:
        v = gaussian_3dconv(v, 5, 1)
    return v


def gather_rays(rays_o_tr, rays_d_tr, viewdirs_tr, rgb_tr, sel):
    sel = np.asarray(sel, dtype=np.int64)
    return rays_o_tr[sel], rays_d_tr[sel], viewdirs_tr[sel], rgb_tr[sel]


class RefSchedule:
    """Learning rates / TV terms / s-terms as the reference's loop leaves them after each iteration.
    cfg_train: reference-style dict (nested tv_terms); cfg_model: dict with optional s_updates."""

    def __init__(self, cfg_train, cfg_model=None):
        self.t = cfg_train
        self.m = dict(cfg_model or {})
        self.tv = dict(cfg_train.get("tv_terms", {}))
        self.s = {}
        self.names = [k[6:] for k in cfg_train if k.startswith("lrate_") and k != "lrate_decay" and cfg_train[k] > 0]
        self._fresh_optimizer(0)

    def _fresh_optimizer(self, global_step):
        # create_optimizer_or_freeze_model, nerf_training.py:9-37
        decay = 0.1 ** (global_step / (self.t["lrate_decay"] * 1000))
        self.lr = {n: self.t["lrate_" + n] * decay for n in self.names}

    @staticmethod
    def _cos(it, warm, warm_min, max_steps, const_warm, min_ratio):
        # nerf_training.py:396-405
        if it < warm:
            return warm_min if const_warm else warm_min + (1 - warm_min) * (it / warm)
        return (1 + math.cos((it - warm) / (max_steps - warm) * math.pi)) * 0.5 * (1 - min_ratio) + min_ratio

    def iteration(self, global_step):
        t = self.t
        if global_step in t.get("pg_scale", []):          # :244-254 (before the step)
            self._fresh_optimizer(0)
        g_ = global_step - 1                               # :386
        if not t.get("cosine_lr", ""):                     # :391-395
            f = 0.1 ** (1 / (t["lrate_decay"] * 1000))
        else:                                              # :396-433
            c = t.get("cosine_lr_cfg", {})
            args = (c.get("warm_up_iters", 0), c.get("warm_up_min_ratio", 1.0), t["N_iters"],
                    c.get("const_warm_up", False), c.get("cos_min_ratio", False))
            pre = 1.0 if global_step == 0 else self._cos(g_ - 1, *args)
            f = self._cos(g_, *args) / pre
        for n in self.lr:
            self.lr[n] = self.lr[n] * f
        for n, fac in t.get("decay_step_module", {}).get(g_, {}).items():
            if n in self.lr:
                self.lr[n] = self.lr[n] * fac
        for k, v in t.get("tv_updates", {}).get(g_, {}).items():
            self.tv[k] = v
        for k, v in self.m.get("s_updates", {}).get(g_, {}).items():
            self.s[k] = v
