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
import itertools

import numpy as np
import torch

from . import _lib
from ._lib import check, lib, ptr, stream

from . import engine  # noqa: F401  (registers the ctypes signatures of include/fgs_b200_train.h)


def extract_fields(bound_min, bound_max, resolution, query_func, N=64):
    """Any scalar field on the resolution^3 lattice between the two corners, queried in N^3 blocks in the reference's
    block order; the result is assembled on the device and copied to the host once."""
    dev = bound_min.device if torch.is_tensor(bound_min) else None
    axes = [torch.linspace(float(bound_min[c]), float(bound_max[c]), resolution, device=dev) for c in range(3)]
    starts = list(range(0, resolution, N))
    field = None
    with torch.no_grad():
        for x0, y0, z0 in itertools.product(starts, starts, starts):
            sub = [ax[o:o + N] for ax, o in zip(axes, (x0, y0, z0))]
            pts = torch.stack(torch.meshgrid(*sub, indexing='ij'), dim=-1).reshape(-1, 3)
            val = query_func(pts).detach().reshape([len(t) for t in sub])
            if field is None:
                field = torch.zeros((resolution,) * 3, dtype=torch.float32, device=val.device)
            field[x0:x0 + val.shape[0], y0:y0 + val.shape[1], z0:z0 + val.shape[2]] = val
    return field.cpu().numpy()


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
