"""Drop-in for the reference's `render_utils_cuda` extension module.

Same function names, argument order/meaning, return lists and error behaviour as the 13 pybind
exports of /root/reference model/cuda/render_utils.cpp:170-184, backed by the sm_100a kernels
in csrc/sampling.cu and csrc/alpha2weight.cu through the C-ABI (include/fgs_b200.h).
Functional ops return freshly allocated CUDA tensors; nothing here computes on the CPU.
"""
import torch

from . import _lib
from ._lib import check, check_input, lib, ptr, stream

_F32 = torch.float32


def _dev(t):
    return t.device


def infer_t_minmax(rays_o, rays_d, xyz_min, xyz_max, near, far):
    """render_utils.cpp:50-58 -> [t_min, t_max]"""
    for t, n in ((rays_o, "rays_o"), (rays_d, "rays_d"), (xyz_min, "xyz_min"), (xyz_max, "xyz_max")):
        check_input(t, n, _F32)
    n_rays = rays_o.size(0)
    t_min = torch.empty(n_rays, dtype=_F32, device=_dev(rays_o))
    t_max = torch.empty(n_rays, dtype=_F32, device=_dev(rays_o))
    with torch.cuda.device(rays_o.device):
        check(lib.fgs_infer_t_minmax(ptr(rays_o), ptr(rays_d), ptr(xyz_min), ptr(xyz_max),
                                     float(near), float(far), n_rays, ptr(t_min), ptr(t_max), stream()))
    return [t_min, t_max]


def infer_n_samples(rays_d, t_min, t_max, stepdist):
    """render_utils.cpp:60-65 -> int64 [N]"""
    for t, n in ((rays_d, "rays_d"), (t_min, "t_min"), (t_max, "t_max")):
        check_input(t, n, _F32)
    n_rays = t_min.size(0)
    out = torch.empty(n_rays, dtype=torch.int64, device=_dev(t_min))
    with torch.cuda.device(t_min.device):
        check(lib.fgs_infer_n_samples(ptr(rays_d), ptr(t_min), ptr(t_max), float(stepdist), n_rays,
                                      ptr(out), stream()))
    return out


def infer_ray_start_dir(rays_o, rays_d, t_min):
    """render_utils.cpp:67-72 -> [rays_start, rays_dir]"""
    for t, n in ((rays_o, "rays_o"), (rays_d, "rays_d"), (t_min, "t_min")):
        check_input(t, n, _F32)
    n_rays = rays_o.size(0)
    start = torch.empty_like(rays_o)
    rdir = torch.empty_like(rays_o)
    with torch.cuda.device(rays_o.device):
        check(lib.fgs_infer_ray_start_dir(ptr(rays_o), ptr(rays_d), ptr(t_min), n_rays, ptr(start),
                                          ptr(rdir), stream()))
    return [start, rdir]


_scratch = {}


def _sample_scratch(device, n_rays):
    need = lib.fgs_sample_scratch_bytes(n_rays)
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = _scratch.get(key)
    if buf is None or buf.numel() < need:
        buf = torch.zeros(max(need, 4096), dtype=torch.uint8, device=device)
        _scratch[key] = buf
    return buf


def sample_pts_on_rays(rays_o, rays_d, xyz_min, xyz_max, near, far, stepdist):
    """render_utils.cpp:74-85 ->
    [ray_pts [T,3] f32, mask_outbbox [T] bool, ray_id [T] i64, step_id [T] i64,
     N_steps [N] i64, t_min [N] f32, t_max [N] f32]
    Two launches (count, fill) with one read-back of total_len in between -- the reference needs
    6 launches, 2 scans and the same read-back (kernel.cu:212)."""
    for t, n in ((rays_o, "rays_o"), (rays_d, "rays_d"), (xyz_min, "xyz_min"), (xyz_max, "xyz_max")):
        check_input(t, n, _F32)
    assert rays_o.dim() == 2 and rays_o.size(1) == 3
    assert rays_d.dim() == 2 and rays_d.size(1) == 3
    dev = _dev(rays_o)
    n_rays = rays_o.size(0)
    with torch.cuda.device(dev):
        st = stream()
        t_min = torch.empty(n_rays, dtype=_F32, device=dev)
        t_max = torch.empty(n_rays, dtype=_F32, device=dev)
        n_steps = torch.empty(n_rays, dtype=torch.int64, device=dev)
        ray_off = torch.empty(n_rays + 1, dtype=torch.int64, device=dev)
        scratch = _sample_scratch(dev, n_rays)
        check(lib.fgs_sample_pts_count(ptr(rays_o), ptr(rays_d), ptr(xyz_min), ptr(xyz_max),
                                       float(near), float(far), float(stepdist), n_rays, ptr(t_min),
                                       ptr(t_max), ptr(n_steps), ptr(ray_off), ptr(scratch), st))
        total_len = int(ray_off[n_rays].item())
        ray_pts = torch.empty((total_len, 3), dtype=_F32, device=dev)
        mask_outbbox = torch.empty(total_len, dtype=torch.bool, device=dev)
        ray_id = torch.empty(total_len, dtype=torch.int64, device=dev)
        step_id = torch.empty(total_len, dtype=torch.int64, device=dev)
        check(lib.fgs_sample_pts_fill(ptr(rays_o), ptr(rays_d), ptr(xyz_min), ptr(xyz_max), ptr(t_min),
                                      ptr(ray_off), float(stepdist), n_rays, total_len, ptr(ray_pts),
                                      ptr(mask_outbbox), ptr(ray_id), ptr(step_id), st))
    return [ray_pts, mask_outbbox, ray_id, step_id, n_steps, t_min, t_max]


def sample_ndc_pts_on_rays(rays_o, rays_d, xyz_min, xyz_max, N_samples):
    """render_utils.cpp:87-97 -> [ray_pts [N,S,3], mask_outbbox [N,S]]"""
    for t, n in ((rays_o, "rays_o"), (rays_d, "rays_d"), (xyz_min, "xyz_min"), (xyz_max, "xyz_max")):
        check_input(t, n, _F32)
    dev = _dev(rays_o)
    n_rays = rays_o.size(0)
    pts = torch.empty((n_rays, N_samples, 3), dtype=_F32, device=dev)
    mask = torch.empty((n_rays, N_samples), dtype=torch.bool, device=dev)
    with torch.cuda.device(dev):
        check(lib.fgs_sample_ndc_pts_on_rays(ptr(rays_o), ptr(rays_d), ptr(xyz_min), ptr(xyz_max),
                                             int(N_samples), n_rays, ptr(pts), ptr(mask), stream()))
    return [pts, mask]


def sample_bg_pts_on_rays(rays_o, rays_d, t_max, bg_preserve, N_samples):
    """render_utils.cpp:99-107 -> ray_pts [N,S,3]"""
    for t, n in ((rays_o, "rays_o"), (rays_d, "rays_d"), (t_max, "t_max")):
        check_input(t, n, _F32)
    dev = _dev(rays_o)
    n_rays = rays_o.size(0)
    pts = torch.empty((n_rays, N_samples, 3), dtype=_F32, device=dev)
    with torch.cuda.device(dev):
        check(lib.fgs_sample_bg_pts_on_rays(ptr(rays_o), ptr(rays_d), ptr(t_max), float(bg_preserve),
                                            int(N_samples), n_rays, ptr(pts), stream()))
    return pts


def maskcache_lookup(world, xyz, xyz2ijk_scale, xyz2ijk_shift):
    """render_utils.cpp:109-118 -> bool [M]"""
    check_input(world, "world", torch.bool)
    check_input(xyz, "xyz", _F32)
    check_input(xyz2ijk_scale, "xyz2ijk_scale", _F32)
    check_input(xyz2ijk_shift, "xyz2ijk_shift", _F32)
    assert world.dim() == 3
    assert xyz.dim() == 2
    assert xyz.size(1) == 3
    n_pts = xyz.size(0)
    out = torch.empty(n_pts, dtype=torch.bool, device=_dev(xyz))
    with torch.cuda.device(xyz.device):
        check(lib.fgs_maskcache_lookup(ptr(world), ptr(xyz), ptr(xyz2ijk_scale), ptr(xyz2ijk_shift),
                                       world.size(0), world.size(1), world.size(2), n_pts, ptr(out),
                                       stream()))
    return out


def raw2alpha(density, shift, interval):
    """render_utils.cpp:120-125 -> [exp_d, alpha]"""
    check_input(density, "density", _F32)
    assert density.dim() == 1
    exp_d = torch.empty_like(density)
    alpha = torch.empty_like(density)
    with torch.cuda.device(density.device):
        check(lib.fgs_raw2alpha(ptr(density), float(shift), float(interval), None, density.size(0),
                                ptr(exp_d), ptr(alpha), stream()))
    return [exp_d, alpha]


def raw2alpha_nonuni(density, shift, interval):
    """render_utils.cpp:127-130 (per-point interval tensor)"""
    check_input(density, "density", _F32)
    check_input(interval, "interval", _F32)
    assert density.dim() == 1
    exp_d = torch.empty_like(density)
    alpha = torch.empty_like(density)
    with torch.cuda.device(density.device):
        check(lib.fgs_raw2alpha(ptr(density), float(shift), 0.0, ptr(interval), density.size(0),
                                ptr(exp_d), ptr(alpha), stream()))
    return [exp_d, alpha]


def raw2alpha_backward(exp, grad_back, interval):
    """render_utils.cpp:132-136"""
    check_input(exp, "exp", _F32)
    check_input(grad_back, "grad_back", _F32)
    grad = torch.empty_like(exp)
    with torch.cuda.device(exp.device):
        check(lib.fgs_raw2alpha_backward(ptr(exp), ptr(grad_back), float(interval), None, exp.size(0),
                                         ptr(grad), stream()))
    return grad


def raw2alpha_nonuni_backward(exp, grad_back, interval):
    """render_utils.cpp:137-140"""
    check_input(exp, "exp", _F32)
    check_input(grad_back, "grad_back", _F32)
    check_input(interval, "interval", _F32)
    grad = torch.empty_like(exp)
    with torch.cuda.device(exp.device):
        check(lib.fgs_raw2alpha_backward(ptr(exp), ptr(grad_back), 0.0, ptr(interval), exp.size(0),
                                         ptr(grad), stream()))
    return grad


def alpha2weight(alpha, ray_id, n_rays):
    """render_utils.cpp:142-148 -> [weight, T, alphainv_last, i_start, i_end]"""
    check_input(alpha, "alpha", _F32)
    check_input(ray_id, "ray_id", torch.int64)
    assert alpha.dim() == 1
    assert ray_id.dim() == 1
    assert alpha.size(0) == ray_id.size(0)
    dev = _dev(alpha)
    n_pts = alpha.size(0)
    weight = torch.empty_like(alpha)
    T = torch.empty_like(alpha)
    alphainv_last = torch.empty(n_rays, dtype=_F32, device=dev)
    seg = torch.empty((2, n_rays), dtype=torch.int64, device=dev)
    i_start, i_end = seg[0], seg[1]
    with torch.cuda.device(dev):
        check(lib.fgs_alpha2weight(ptr(alpha), ptr(ray_id), n_pts, int(n_rays), ptr(weight), ptr(T),
                                   ptr(alphainv_last), ptr(i_start), ptr(i_end), stream()))
    return [weight, T, alphainv_last, i_start, i_end]


def alpha2weight_backward(alpha, weight, T, alphainv_last, i_start, i_end, n_rays, grad_weights,
                          grad_last):
    """render_utils.cpp:150-167 -> grad_alpha"""
    for t, n in ((alpha, "alpha"), (weight, "weight"), (T, "T"), (alphainv_last, "alphainv_last"),
                 (grad_weights, "grad_weights"), (grad_last, "grad_last")):
        check_input(t, n, _F32)
    check_input(i_start, "i_start", torch.int64)
    check_input(i_end, "i_end", torch.int64)
    grad = torch.empty_like(alpha)
    with torch.cuda.device(alpha.device):
        check(lib.fgs_alpha2weight_backward(ptr(alpha), ptr(weight), ptr(T), ptr(alphainv_last),
                                            ptr(i_start), ptr(i_end), int(n_rays), alpha.size(0),
                                            ptr(grad_weights), ptr(grad_last), ptr(grad), stream()))
    return grad
