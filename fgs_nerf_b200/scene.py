"""Deterministic synthetic scenes and ray batches of the benchmark shapes.

Everything is generated on the CPU with fixed seeds so the GPU path, the oracle and the golden
fixtures all see identical inputs:
  * cube bbox [-1,1]^3, G^3 voxels, voxel_size by the reference's formula (model/nerf.py:298-303),
  * SDF grid  ||p|| - 0.6 + 0.01*N(0,1) (seed 777) then one 3^3 box blur (ball-init style,
    model/nerf.py:77-82),  k0 = 0.1*N(0,1) (seed 778),
  * mask grid (the MaskCache of model/nerf.py:1192-1209): 120^3 over the bbox x1.2, 1e-3 inside
    the shell | ||p||-0.6 | < 0.15, max-pooled k=3 like nerf.py:1199, threshold 1e-3,
  * MLPs with nn.Linear default init under seed 779, widths of config/shiny_blender.py:221-250,
  * rays of a blender-style pinhole camera (model/nerf_ray.py:8-38, lib/load_blender.py:29-34):
    800x800, r=4, elevation -30 deg, 40-view ring; batch = first N of a seeded permutation.
"""
import math
from dataclasses import dataclass, field

import numpy as np
import torch
import torch.nn.functional as F

FINE_MODEL = dict(  # config/shiny_blender.py:221-250 (same in dtu.py / smart_car.py)
    k0_dim=12, rgbnet_width=256, rgbnet_depth=4, refnet_width=256, refnet_depth=4,
    posbase_pe=5, viewbase_pe=3, refbase_pe=8,
    grad_feat=(0.5, 1.0, 1.5, 2.0), sdf_feat=(0.5, 1.0, 1.5, 2.0),
    s_ratio=50, s_start=0.05, step_start=0, stepsize=0.5,
    fast_color_thres=1e-4, mask_cache_thres=1e-3,
)
FINE_TRAIN = dict(  # config/shiny_blender.py:180-219
    lrate_k0=0.1, lrate_sdf=0.005, lrate_rgbnet=1e-3, lrate_refnet=1e-3,
    weight_main=1.0, weight_entropy_last=1e-3, weight_orientation=1e-4, sigmoid_rgb_loss=0.02,
    weight_tv_density=0.01, weight_tv_k0=0.0, tv_every=3, tv_from=0, tv_end=30000,
    tv_dense_before=20000, sdf_tv=0.1, smooth_grad_tv=0.05, N_iters=20000,
)


def mlp_dims(k0_dim=12, posbase_pe=5, viewbase_pe=3, refbase_pe=8, n_feat=4, width=256):
    """rgbnet_dim / refnet_dim of model/nerf.py:108-116 (fine stage)."""
    rgbnet_dim = (3 + 3 * posbase_pe * 2) + k0_dim + 3 + n_feat * 3 + n_feat * 6 + 1 + (3 + 3 * viewbase_pe * 2)
    refnet_dim = (3 + 3 * refbase_pe * 2) + width
    return rgbnet_dim, refnet_dim


@dataclass
class Scene:
    G: int
    xyz_min: torch.Tensor
    xyz_max: torch.Tensor
    voxel_size: float
    sdf: torch.Tensor            # [1,1,G,G,G]
    k0: torch.Tensor             # [1,12,G,G,G] (logical NCDHW)
    mask_grid: torch.Tensor      # [1,1,Gm,Gm,Gm] (already max-pooled)
    mask_xyz_min: torch.Tensor
    mask_xyz_max: torch.Tensor
    mask_thres: float
    rgbnet: list = field(default_factory=list)   # [(W [out,in], b [out])] x4
    refnet: list = field(default_factory=list)
    cfg: dict = field(default_factory=dict)


def make_mlps(seed=779, cfg=FINE_MODEL):
    """nn.Linear default init; construction order refnet then rgbnet as model/nerf.py:125-142."""
    rgb_dim, ref_dim = mlp_dims(cfg["k0_dim"], cfg["posbase_pe"], cfg["viewbase_pe"], cfg["refbase_pe"],
                                len(cfg["grad_feat"]), cfg["rgbnet_width"])
    g = torch.Generator().manual_seed(seed)

    def lin(i, o):
        bound = 1.0 / math.sqrt(i)  # kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(in), 1/sqrt(in))
        w = (torch.rand(o, i, generator=g) * 2 - 1) * bound
        b = (torch.rand(o, generator=g) * 2 - 1) * bound
        return w.float().contiguous(), b.float().contiguous()

    W = cfg["refnet_width"]
    refnet = [lin(ref_dim, W)] + [lin(W, W) for _ in range(cfg["refnet_depth"] - 2)] + [lin(W, 3)]
    W = cfg["rgbnet_width"]
    rgbnet = [lin(rgb_dim, W)] + [lin(W, W) for _ in range(cfg["rgbnet_depth"] - 2)] + [lin(W, W)]
    return rgbnet, refnet


def make_scene(G=160, mask_G=120, seed=777, cfg=FINE_MODEL, with_mlps=True):
    xyz_min = torch.tensor([-1.0, -1.0, -1.0])
    xyz_max = torch.tensor([1.0, 1.0, 1.0])
    # model/nerf.py:301 (float32 arithmetic); world_size is set explicitly to G^3
    voxel_size = float(((xyz_max - xyz_min).prod() / float(G ** 3)).pow(1 / 3))
    lin = torch.linspace(-1.0, 1.0, G)
    x, y, z = torch.meshgrid(lin, lin, lin, indexing="ij")
    g = torch.Generator().manual_seed(seed)
    sdf = (x * x + y * y + z * z).sqrt() - 0.6 + 0.01 * torch.randn(G, G, G, generator=g)
    sdf = F.avg_pool3d(F.pad(sdf[None, None], (1, 1, 1, 1, 1, 1), mode="replicate"), 3, stride=1)
    g = torch.Generator().manual_seed(seed + 1)
    k0 = 0.1 * torch.randn(1, cfg["k0_dim"], G, G, G, generator=g)
    mlin = torch.linspace(-1.2, 1.2, mask_G)
    mx, my, mz = torch.meshgrid(mlin, mlin, mlin, indexing="ij")
    shell = (((mx * mx + my * my + mz * mz).sqrt() - 0.6).abs() < 0.15).float() * 1e-3
    mask_grid = F.max_pool3d(shell[None, None], kernel_size=3, padding=1, stride=1)
    sc = Scene(G=G, xyz_min=xyz_min, xyz_max=xyz_max, voxel_size=voxel_size,
               sdf=sdf.float().contiguous(), k0=k0.float().contiguous(),
               mask_grid=mask_grid.float().contiguous(),
               mask_xyz_min=torch.tensor([-1.2, -1.2, -1.2]), mask_xyz_max=torch.tensor([1.2, 1.2, 1.2]),
               mask_thres=cfg["mask_cache_thres"], cfg=dict(cfg))
    if with_mlps:
        sc.rgbnet, sc.refnet = make_mlps(seed + 2, cfg)
    return sc


def _pose_spherical(theta_deg, phi_deg, radius):
    """lib/load_blender.py:11-34, float64 then float32."""
    th, ph = math.radians(theta_deg), math.radians(phi_deg)
    trans = np.eye(4); trans[2, 3] = radius
    rphi = np.array([[1, 0, 0, 0], [0, math.cos(ph), -math.sin(ph), 0], [0, math.sin(ph), math.cos(ph), 0], [0, 0, 0, 1.0]])
    rth = np.array([[math.cos(th), 0, -math.sin(th), 0], [0, 1, 0, 0], [math.sin(th), 0, math.cos(th), 0], [0, 0, 0, 1.0]])
    flip = np.array([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1.0]])
    return (flip @ rth @ rphi @ trans).astype(np.float32)


def make_rays(N=8192, H=800, W=800, n_views=40, seed=777, inverse_y=False, radius=4.0, elev=-30.0,
              camera_angle_x=0.6911, offset=0):
    """First N (from `offset`) entries of a seeded permutation over all n_views*H*W pixels; rays are
    generated only for the selected pixels (pixel-centre mode of model/nerf_ray.py:17-18).
    Returns float32 CPU tensors rays_o, rays_d, viewdirs [N,3] and target rgb [N,3] (seed+3)."""
    focal = 0.5 * W / math.tan(0.5 * camera_angle_x)
    total = n_views * H * W
    perm = np.random.RandomState(seed).permutation(total)[offset:offset + N]
    view = perm // (H * W)
    pix = perm % (H * W)
    row = (pix // W).astype(np.float32) + 0.5   # j
    col = (pix % W).astype(np.float32) + 0.5    # i
    cx, cy = np.float32(0.5 * W), np.float32(0.5 * H)
    f = np.float32(focal)
    if inverse_y:
        dirs = np.stack([(col - cx) / f, (row - cy) / f, np.ones_like(col)], -1)
    else:
        dirs = np.stack([(col - cx) / f, -(row - cy) / f, -np.ones_like(col)], -1)
    poses = np.stack([_pose_spherical(t, elev, radius) for t in np.linspace(-180, 180, n_views + 1)[:-1]])
    R = poses[view, :3, :3]                       # [N,3,3]
    if inverse_y:
        # OpenCV-convention cameras (DTU, model/nerf_ray.py:24-25: x right, y down, z forward): the same
        # viewpoints expressed in that convention, c2w = c2w_gl @ diag(1, -1, -1)
        R = R * np.array([1.0, -1.0, -1.0], dtype=np.float32)[None, None, :]
    rays_d = np.einsum("nij,nj->ni", R, dirs.astype(np.float32)).astype(np.float32)
    rays_o = poses[view, :3, 3].astype(np.float32)
    viewdirs = rays_d / np.linalg.norm(rays_d, axis=-1, keepdims=True)
    target = np.random.RandomState(seed + 3).uniform(0, 1, size=(N, 3)).astype(np.float32)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a.astype(np.float32)))
    return t(rays_o), t(rays_d), t(viewdirs), t(target)


def s_val_at(global_step, cfg=FINE_MODEL):
    """model/nerf.py:514 (python floats)."""
    return 1.0 / (global_step + cfg["s_ratio"] / cfg["s_start"] - cfg["step_start"]) * cfg["s_ratio"]
