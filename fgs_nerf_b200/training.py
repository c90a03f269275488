"""Host side of the fine-stage training loop (model/nerf_training.py:237-460) around
FineStepEngine.train_step: batch sampling, learning-rate / TV / s_val schedules, progressive grid
scaling -- SURVEY.md 8(f) rows 2 and 3.

What changes against the reference loop, and what does not:
  * the batch indices follow the reference's generator exactly (one numpy permutation per epoch,
    nerf_ray.py:252-259) but live on the DEVICE: one upload per epoch instead of a CPU index tensor
    per step, and the four `x_tr[sel_i]` selections are one launch (fgs_gather_rays);
  * nothing crosses to the host inside an iteration (the reference reads psnr with .item() and five
    weight statistics with .cpu() every step, nerf_training.py:309,375-384): the loss terms of every
    step are kept in a device ring and read when stats() is called (every i_print steps);
  * the learning rates are python floats multiplied step by step in the reference's order (plain
    exponential or cosine ratio, then decay_step_module), the optimiser is re-created with base rates
    after a pg_scale step (:244-254), tv_updates / s_updates take effect from the next iteration
    (:416-428) -- all as the reference does, quirks included.
Not provided: the `patch` / `random` samplers, voxel_inc (coarse stage only), reset_iter.
"""
import math

import numpy as np
import torch

from . import dist as fdist
from .engine import gather_rays


def batch_indices_generator(N, BS, rng=None):
    """nerf_ray.batch_indices_generator (model/nerf_ray.py:252-259) on numpy int64 arrays: a fresh
    permutation whenever the next batch would run past the end (the tail of an epoch is dropped).
    rng=None uses numpy's global state, like the reference."""
    rng = np.random if rng is None else rng
    idx, top = rng.permutation(N), 0
    while True:
        if top + BS > N:
            idx, top = rng.permutation(N), 0
        yield idx[top:top + BS]
        top += BS


class DeviceBatchSampler:
    """The same index stream as batch_indices_generator, as views of a device-resident permutation."""

    def __init__(self, N, BS, device, rng=None):
        if BS > N:
            raise ValueError(f"batch size {BS} exceeds the number of training rays {N}")
        self.N, self.BS, self.device = int(N), int(BS), torch.device(device)
        self.rng = np.random if rng is None else rng
        self._host = torch.empty(self.N, dtype=torch.int64).pin_memory() if self.device.type == "cuda" \
            else torch.empty(self.N, dtype=torch.int64)
        self._perm = torch.empty(self.N, dtype=torch.int64, device=self.device)
        self._uploaded = None
        self._new_epoch()

    def _new_epoch(self):
        if self._uploaded is not None:
            self._uploaded.synchronize()          # the pinned buffer is about to be overwritten
        self._host.copy_(torch.from_numpy(np.ascontiguousarray(self.rng.permutation(self.N), dtype=np.int64)))
        self._perm.copy_(self._host, non_blocking=True)
        if self.device.type == "cuda":
            self._uploaded = torch.cuda.Event()
            self._uploaded.record(torch.cuda.current_stream(self.device))
        self.top = 0

    def __iter__(self):
        return self

    def __next__(self):
        if self.top + self.BS > self.N:
            self._new_epoch()
        sel = self._perm[self.top:self.top + self.BS]
        self.top += self.BS
        return sel


def cosine_lr_func(iter, warm_up_iters, warm_up_min_ratio, max_steps, const_warm_up=False, min_ratio=0):
    """model/nerf_training.py:396-405."""
    if iter < warm_up_iters:
        if not const_warm_up:
            lr = warm_up_min_ratio + (1 - warm_up_min_ratio) * (iter / warm_up_iters)
        else:
            lr = warm_up_min_ratio
    else:
        lr = (1 + math.cos((iter - warm_up_iters) / (max_steps - warm_up_iters) * math.pi)) * 0.5 * (
            1 - min_ratio) + min_ratio
    return lr


def lr_decay_factor(cfg_train, global_step):
    """Factor every param group's lr is multiplied by at the end of iteration `global_step`
    (model/nerf_training.py:386-433)."""
    global_step_ = global_step - 1
    if not cfg_train.get("cosine_lr", ""):
        decay_steps = cfg_train["lrate_decay"] * 1000
        return 0.1 ** (1 / decay_steps)
    cc = cfg_train.get("cosine_lr_cfg", {})
    warm_up_iters = cc.get("warm_up_iters", 0)
    warm_up_min_ratio = cc.get("warm_up_min_ratio", 1.0)
    const_warm_up = cc.get("const_warm_up", False)
    cos_min_ratio = cc.get("cos_min_ratio", False)
    N_iters = cfg_train["N_iters"]
    if global_step == 0:
        pre = 1.0
    else:
        pre = cosine_lr_func(global_step_ - 1, warm_up_iters, warm_up_min_ratio, N_iters, const_warm_up, cos_min_ratio)
    pos = cosine_lr_func(global_step_, warm_up_iters, warm_up_min_ratio, N_iters, const_warm_up, cos_min_ratio)
    return pos / pre


def end_of_iteration_updates(t, lr, global_step, engine_train_cfg=None, model_cfg=None):
    """The schedule block that closes an iteration (model/nerf_training.py:386-433), applied after
    optimizer.step(): every group's lr times the decay factor, then decay_step_module, tv_updates and s_updates
    keyed on global_step - 1.  `lr` (group name -> python float), `t` (flat train cfg) and the two optional dicts
    (the engine's copies) are updated in place."""
    f = lr_decay_factor(t, global_step)
    for k in lr:
        lr[k] = lr[k] * f
    gs_ = global_step - 1
    for name, fac in t.get("decay_step_module", {}).get(gs_, {}).items():
        if name in lr:
            lr[name] = lr[name] * fac
    for term, value in t.get("tv_updates", {}).get(gs_, {}).items():
        t[term] = value
        if engine_train_cfg is not None:
            engine_train_cfg[term] = value
    # s_updates / smooth_updates live in the MODEL config in the reference (getattr(cfg_model, ...), :446-456); the
    # train cfg is accepted too
    s_updates = dict(t.get("s_updates", {}))
    if model_cfg is not None:
        s_updates.update(model_cfg.get("s_updates", {}))      # (smooth_updates: CoarseTrainer.step, after this block)
    for term, value in s_updates.get(gs_, {}).items():
        if model_cfg is not None:
            model_cfg[term] = value


def flatten_train_cfg(cfg_train):
    """Reference-style fine_train dict (config/shiny_blender.py:180-219, nested tv_terms) -> the flat dict
    FineStepEngine reads.  Unsupported switches of the reference loop fail loudly."""
    t = dict(cfg_train)
    for k, v in dict(t.pop("tv_terms", {})).items():
        t[k] = v
    t.setdefault("sdf_tv", 0.0)
    t.setdefault("smooth_grad_tv", 0.0)
    if t.get("weight_rgbper", 0) > 0:
        raise NotImplementedError("weight_rgbper > 0 (coarse-stage loss) is not on the fine-stage path")
    if t.get("ori_tv", False) or t.get("weight_tv_k0", 0) > 0 or t.get("grad_norm", 0) > 0 or t.get("grad_tv", 0) > 0:
        raise NotImplementedError("ori_tv / weight_tv_k0 / grad_norm / grad_tv are 0 in every fine-stage config")
    if t.get("voxel_inc", False) or t.get("pervoxel_lr", False) or t.get("reset_iter"):
        raise NotImplementedError("voxel_inc / pervoxel_lr / reset_iter belong to the coarse stages")
    if t.get("smooth_updates"):
        raise NotImplementedError("smooth_updates (init_smooth_conv, coarse stages) is not on the fine-stage path")
    if t.get("ray_sampler", "in_maskcache") not in ("flatten", "in_maskcache"):
        raise NotImplementedError(f"ray_sampler {t['ray_sampler']!r}")
    return t


def flatten_coarse_train_cfg(cfg_train):
    """Reference-style coarse_train dict (config/shiny_blender.py:103-146) -> the flat dict CoarseStepEngine reads:
    weight_rgbper and ori_tv belong to this stage; what the coarse kernels do not cover fails loudly."""
    t = dict(cfg_train)
    for k, v in dict(t.pop("tv_terms", {})).items():
        t[k] = v
    t.setdefault("sdf_tv", 0.0)
    t.setdefault("smooth_grad_tv", 0.0)
    if t.get("weight_tv_k0", 0) > 0 or t.get("grad_norm", 0) > 0 or t.get("grad_tv", 0) > 0:
        raise NotImplementedError("weight_tv_k0 / grad_norm / grad_tv are 0 in every shipped coarse config")
    if t.get("pervoxel_lr", False):
        raise NotImplementedError("pervoxel_lr is off in every shipped config")
    if t.get("ray_sampler", "in_maskcache") not in ("flatten", "in_maskcache", "random"):
        raise NotImplementedError(f"ray_sampler {t['ray_sampler']!r}")
    return t


class FineTrainer:
    """scene_rep_reconstruction's iteration loop for stage 'fine' (model/nerf_training.py:237-460).

    engine        FineStepEngine for the PER-RANK batch (cfg_train['N_rand'] / world_size rays)
    *_tr          flattened training rays [n_total,3] on the engine's device (what get_training_rays_* return)
    cfg_train     reference-style dict (tv_terms nested or flat)
    start         the loop runs global_step = start + 1 ... (checkpoint resume: nerf_training.py:120-133)
    reset_optimizer  None (default): a fresh MaskedAdam at the base learning rates -- create_optimizer_or_freeze_model(...,
                  global_step=0), nerf_training.py:117 -- UNLESS the engine already carries optimiser state (loaded by
                  checkpoint.engine_from_checkpoint = utils.load_checkpoint's optimizer.load_state_dict, utils.py:49-56),
                  which is then kept: moments, step count and the per-group learning rates.  True forces the fresh
                  optimiser (the reference's --no_reload_optimizer), False never touches the engine's.
    rng           numpy RandomState of the batch index stream (None: numpy's global state, which the reference seeds
                  in run.py).  Data parallel: every rank must run the SAME stream (same seed) -- each keeps its
                  contiguous slice of the global batch (dist.ray_shard), the last rank owns the batch's last ray.
    """

    _flatten = staticmethod(flatten_train_cfg)
    _stage = "fine"

    def __init__(self, engine, rays_o_tr, rays_d_tr, viewdirs_tr, rgb_tr, cfg_train=None, start=0, near=2.0, bg=1.0,
                 rng=None, ring=256, reset_optimizer=None):
        self.engine = engine
        self.cfg_train = self._flatten(cfg_train if cfg_train is not None else engine.train_cfg)
        engine.train_cfg.update({k: v for k, v in self.cfg_train.items() if k in engine.train_cfg or k in
                                 ("sdf_tv", "smooth_grad_tv", "lrate_decay")})
        self.world_size, self.rank = engine.world_size, engine.rank
        self.N_rand = int(self.cfg_train.get("N_rand", engine.n_rays * self.world_size))
        lo, hi = fdist.ray_shard(self.N_rand, self.world_size, self.rank)
        if hi - lo != engine.n_rays:
            raise ValueError(f"engine.n_rays = {engine.n_rays} but this rank's share of N_rand = {self.N_rand} is {hi - lo}")
        self._shard = (lo, hi)
        self.rays = (rays_o_tr, rays_d_tr, viewdirs_tr, rgb_tr)
        dev = engine.device
        self.sampler = DeviceBatchSampler(rays_o_tr.shape[0], self.N_rand, dev, rng)
        self._batch = tuple(torch.empty((engine.n_rays, 3), dtype=torch.float32, device=dev) for _ in range(4))
        self._bad = torch.zeros(1, dtype=torch.int32, device=dev)
        self.near, self.bg = float(near), float(bg)
        self.global_step = int(start)
        self._ring = torch.zeros((int(ring), engine.ws["loss"].numel()), dtype=torch.float64, device=dev)
        self._since = 0                      # iterations since the last stats()
        if reset_optimizer is None:
            reset_optimizer = not (engine.adam and engine.adam_step > 0)
        if reset_optimizer:
            engine.reset_optimizer(0)            # the reference always builds the optimiser with global_step=0

    # ---------------------------------------------------------------------------------
    def step(self):
        """One iteration; enqueues only (plus the data-parallel brick count).  Returns global_step."""
        self.global_step += 1
        gs = self.global_step
        t = self.cfg_train
        if gs in t.get("pg_scale", []):
            # model.scale_volume_grid(num_voxels * scale_ratio) + a fresh optimiser (nerf_training.py:244-254)
            self.engine = self.engine.scale_volume_grid(self.engine.num_voxels * t["scale_ratio"])
            self._after_rescale(gs)
        self._before_iteration(gs)
        eng = self.engine
        sel = next(self.sampler)
        lo, hi = self._shard
        ro, rd, vd, tgt = gather_rays(*self.rays, sel[lo:hi], out=self._batch, bad=self._bad)
        eng.train_step(ro, rd, vd, tgt, gs, near=self.near, bg=self.bg, n_rays_global=self.N_rand,
                       has_last_ray=(self.rank == self.world_size - 1))
        self._ring[self._since % self._ring.shape[0]].copy_(eng.ws["loss"])
        self._since += 1
        end_of_iteration_updates(t, eng.lr, gs, eng.train_cfg, eng.cfg)
        self._after_updates(gs)
        return gs

    def _after_rescale(self, gs):
        pass

    def _before_iteration(self, gs):
        pass

    def _after_updates(self, gs):
        if self.engine.cfg.get("smooth_updates"):
            raise NotImplementedError("smooth_updates (init_smooth_conv) belongs to the coarse stages: use CoarseTrainer")

    def run(self, n_steps, i_print=0, log=print):
        for _ in range(int(n_steps)):
            gs = self.step()
            if i_print and gs % i_print == 0:
                s = self.stats()
                log(f"scene_rep_reconstruction ({self._stage}): iter {gs:6d} / Loss: {s['loss']:.9f} / PSNR: {s['psnr']:5.2f} / "
                    f"s_val: {self.engine.s_val(gs):5.2g}")
        if not i_print:
            self.engine.check_overflow()               # nobody called stats(): still never return silently corrupted
        return self.global_step

    def stats(self, reduce=True):
        """Synchronises.  Loss terms of the (at most `ring`) iterations since the last call: their mean PSNR (the
        reference averages psnr_lst the same way), the last total loss; also checks the batch indices and the
        workspace capacity."""
        R = self._ring.shape[0]
        since, self._since = self._since, 0
        n = min(since, R)
        ring = self._ring.clone()
        bad = self._bad.clone()
        if reduce and self.world_size > 1:
            import torch.distributed as dist
            dist.all_reduce(ring)                      # per-rank terms are already normalised by the global ray count
            dist.all_reduce(bad, op=dist.ReduceOp.MAX)
        ring = ring.cpu().numpy()
        if int(bad.item()):
            raise RuntimeError("batch index outside the training rays")
        self.engine.check_overflow()                   # sticky over ALL iterations since the last check; raises
        steps = list(range(self.global_step - n + 1, self.global_step + 1))
        if n == 0:
            return dict(steps=steps, loss=float("nan"), psnr=float("nan"), terms=ring[:0], lr=dict(self.engine.lr))
        rows = ring[[(since - n + i) % R for i in range(n)]]          # oldest -> newest
        mse = np.maximum(rows[:, 0], 1e-30)
        psnr = float(np.mean(-10.0 * np.log10(mse)))  # utils.mse2psnr on weight_main * mse (nerf_training.py:308-309)
        return dict(steps=steps, loss=float(rows[-1].sum()), psnr=psnr, terms=rows, lr=dict(self.engine.lr))


class CoarseTrainer(FineTrainer):
    """scene_rep_reconstruction's iteration loop for the stages that run forward_coarse (model/nerf_training.py:237-460 with
    `coarse_train` / `coarse_model` or `geometry_searching` / `geometry_searching_model`): the same loop on a
    coarse.CoarseStepEngine -- batch sampler, one-launch gather, lr decay, decay_step_module, tv_updates, s_updates,
    pg_scale (scale_volume_grid + a fresh optimiser) -- plus the coarse model's `smooth_updates` (init_smooth_conv) and,
    for geometry_searching, Incremental Voxel Training (`voxel_inc`: the box of `set_inc_mask` grows over `inc_steps`
    iterations, :201-214, :286-293) and `reset_iter` (a fresh refnet and the near-camera mask at a pg_scale step,
    :249-252; pass the training cameras' origins as `cam_o`).  `ray_sampler='random'` draws pixels uniformly, which on
    flattened rays is what the batch sampler's permutation stream does without replacement.  Single GPU."""
    _flatten = staticmethod(flatten_coarse_train_cfg)
    _stage = "coarse"

    def __init__(self, engine, *rays, cam_o=None, maskout_near_cam_vox=True, **kw):
        super().__init__(engine, *rays, **kw)
        self._stage = engine.stage
        self.cam_o, self.maskout = cam_o, maskout_near_cam_vox
        t = self.cfg_train
        if t.get("voxel_inc", False):
            mid = [t["x_mid"], t["y_mid"], t["z_mid"]]
            r = [t["x_init_ratio"], t["y_init_ratio"], t["z_init_ratio"]]
            self._inc_lo0 = torch.tensor([m - q * m for m, q in zip(mid, r)])            # nerf_training.py:205-214
            self._inc_hi0 = torch.tensor([m + q * (1 - m) for m, q in zip(mid, r)])

    def _after_rescale(self, gs):
        if gs in self.cfg_train.get("reset_iter", []):                 # nerf_training.py:249-252
            self.engine.reset_refnet()
            if self.maskout and self.cam_o is not None:
                self.engine.maskout_near_cam_vox(self.cam_o, self.near)

    def _before_iteration(self, gs):
        t = self.cfg_train
        if t.get("voxel_inc", False):                                  # nerf_training.py:286-293
            if gs <= t["inc_steps"] or self.engine.inc_mask is None:   # (None: a pg_scale step has just replaced the engine)
                w = min(gs * 1.0 / t["inc_steps"], 1.0)
                self.engine.set_inc_mask(self._inc_lo0 - w * self._inc_lo0, self._inc_hi0 + w * (1 - self._inc_hi0))
        elif self.engine.inc_mask is not None:
            self.engine.unset_inc_mask()

    def _after_updates(self, gs):
        upd = self.engine.cfg.get("smooth_updates", {}).get(gs - 1)
        if upd:
            self.engine.init_smooth_conv(**upd)                        # model/nerf_training.py:452-456
