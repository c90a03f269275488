"""Drop-in for the field-query half of the reference's mesh extraction (model/extract_geometry.py):

  extract_fields(bound_min, bound_max, resolution, query_func, N=64) -> np.float32 [res,res,res]
      same signature and block order as the reference (:5-19), for any query_func; the blocks stay on the device
      and cross to the host ONCE (the reference copies each 64^3 block back as it goes).
  extract_sdf_field(sdf_grid, xyz_min, xyz_max, bound_min, bound_max, resolution) -> device tensor
      the query nerf.extract_geometry actually makes (model/nerf.py:1157-1170: grid_sampler(pts, -sdf_grid)) as one
      launch of fgs_grid_sample_lattice: the lattice points are never materialised.
  extract_geometry(...)   marching cubes itself is the CPU library `mcubes` (out of scope);
      it is called when importable, like the reference does.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check, lib, ptr, stream

from . import engine  # noqa: F401  (registers the ctypes signatures of include/fgs_b200_train.h)


def extract_fields(bound_min, bound_max, resolution, query_func, N=64):
    dev = bound_min.device if torch.is_tensor(bound_min) else None
    X = torch.linspace(float(bound_min[0]), float(bound_max[0]), resolution, device=dev).split(N)
    Y = torch.linspace(float(bound_min[1]), float(bound_max[1]), resolution, device=dev).split(N)
    Z = torch.linspace(float(bound_min[2]), float(bound_max[2]), resolution, device=dev).split(N)
    u = None
    with torch.no_grad():
        for xi, xs in enumerate(X):
            for yi, ys in enumerate(Y):
                for zi, zs in enumerate(Z):
                    xx, yy, zz = torch.meshgrid(xs, ys, zs, indexing='ij')
                    pts = torch.stack([xx.reshape(-1), yy.reshape(-1), zz.reshape(-1)], dim=-1)
                    val = query_func(pts).reshape(len(xs), len(ys), len(zs)).detach()
                    if u is None:
                        u = torch.zeros([resolution] * 3, dtype=torch.float32, device=val.device)
                    u[xi * N: xi * N + len(xs), yi * N: yi * N + len(ys), zi * N: zi * N + len(zs)] = val
    return u.cpu().numpy()


@torch.no_grad()
def extract_sdf_field(sdf_grid, xyz_min, xyz_max, bound_min, bound_max, resolution, sign=-1.0):
    """sign * trilinear(sdf_grid [1,1,X,Y,Z]) on the resolution^3 lattice linspace(bound_min, bound_max): the `u` of
    the reference's extract_fields for nerf.extract_geometry's query_func, as a device tensor."""
    _lib.check_input(sdf_grid, "sdf_grid", torch.float32)
    if sdf_grid.dim() != 5 or sdf_grid.shape[0] != 1 or sdf_grid.shape[1] != 1:
        raise RuntimeError(f"sdf_grid must be [1,1,X,Y,Z], got {tuple(sdf_grid.shape)}")
    dev = sdf_grid.device
    X, Y, Z = (int(v) for v in sdf_grid.shape[2:])
    res = int(resolution)
    # torch.linspace on the CPU, as the reference builds it (its bound_min / bound_max are CPU tensors there)
    ax = [torch.linspace(float(bound_min[c]), float(bound_max[c]), res).to(dev) for c in range(3)]
    mn = (C.c_float * 3)(*[float(v) for v in xyz_min])
    mx = (C.c_float * 3)(*[float(v) for v in xyz_max])
    out = torch.empty((res, res, res), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib.fgs_grid_sample_lattice(ptr(sdf_grid), X, Y, Z, mn, mx, ptr(ax[0]), res, ptr(ax[1]), res, ptr(ax[2]), res,
                                          float(sign), ptr(out), stream()))
    return out


def extract_geometry(bound_min, bound_max, resolution, threshold, query_func=None, N=64, field=None):
    """model/extract_geometry.py:21-29.  Pass `field` (from extract_sdf_field) to skip the blockwise query."""
    try:
        import mcubes
    except ImportError as e:
        raise ImportError("marching cubes needs the CPU library `mcubes` (not part of fgs_b200)") from e
    u = field.cpu().numpy() if field is not None else extract_fields(bound_min, bound_max, resolution, query_func, N)
    vertices, triangles = mcubes.marching_cubes(u, threshold)
    b_max_np = np.asarray([float(v) for v in bound_max], dtype=np.float32)
    b_min_np = np.asarray([float(v) for v in bound_min], dtype=np.float32)
    vertices = vertices / (resolution - 1.0) * (b_max_np - b_min_np)[None, :] + b_min_np[None, :]
    return vertices, triangles
