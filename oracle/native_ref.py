"""TEST INFRASTRUCTURE ONLY -- numpy front-end of oracle/fgs_oracle.c (CPU restatement of the
reference's native ops).  Import only from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this.

Function names and return conventions mirror the reference's pybind modules
(model/cuda/render_utils.cpp:170-184, adam_upd.cpp:79-86, total_variation.cpp:29-32) so parity
tests read like calls into the reference.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "fgs_oracle.c")
SO = os.path.join(HERE, "libfgs_oracle.so")


def build(force=False):
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(SRC):
        # -ffp-contract=off: every rounding in the C text is a rounding in the binary
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
                               "-o", SO, SRC, "-lm"])
    return SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.oracle_sample_pts_on_rays.restype = C.c_int64
    return _lib


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def infer_t_minmax(rays_o, rays_d, xyz_min, xyz_max, near, far):
    rays_o, rays_d, xyz_min, xyz_max = map(_f, (rays_o, rays_d, xyz_min, xyz_max))
    n = rays_o.shape[0]
    t_min = np.empty(n, np.float32)
    t_max = np.empty(n, np.float32)
    lib().oracle_infer_t_minmax(_p(rays_o), _p(rays_d), _p(xyz_min), _p(xyz_max), C.c_float(near),
                                C.c_float(far), n, _p(t_min), _p(t_max))
    return [t_min, t_max]


def infer_n_samples(rays_d, t_min, t_max, stepdist):
    rays_d, t_min, t_max = map(_f, (rays_d, t_min, t_max))
    n = t_min.shape[0]
    out = np.empty(n, np.int64)
    lib().oracle_infer_n_samples(_p(rays_d), _p(t_min), _p(t_max), C.c_float(stepdist), n, _p(out))
    return out


def infer_ray_start_dir(rays_o, rays_d, t_min):
    rays_o, rays_d, t_min = map(_f, (rays_o, rays_d, t_min))
    n = rays_o.shape[0]
    start = np.empty_like(rays_o)
    rdir = np.empty_like(rays_o)
    lib().oracle_infer_ray_start_dir(_p(rays_o), _p(rays_d), _p(t_min), n, _p(start), _p(rdir))
    return [start, rdir]


def sample_pts_on_rays(rays_o, rays_d, xyz_min, xyz_max, near, far, stepdist):
    rays_o, rays_d, xyz_min, xyz_max = map(_f, (rays_o, rays_d, xyz_min, xyz_max))
    n = rays_o.shape[0]
    t_min = np.empty(n, np.float32)
    t_max = np.empty(n, np.float32)
    n_steps = np.empty(n, np.int64)
    args = (_p(rays_o), _p(rays_d), _p(xyz_min), _p(xyz_max), C.c_float(near), C.c_float(far),
            C.c_float(stepdist), n, _p(t_min), _p(t_max), _p(n_steps))
    total = lib().oracle_sample_pts_on_rays(*args, None, None, None, None)
    ray_pts = np.empty((total, 3), np.float32)
    mask = np.empty(total, np.uint8)
    ray_id = np.empty(total, np.int64)
    step_id = np.empty(total, np.int64)
    if total > 0:
        lib().oracle_sample_pts_on_rays(*args, _p(ray_pts), _p(mask), _p(ray_id), _p(step_id))
    return [ray_pts, mask.astype(bool), ray_id, step_id, n_steps, t_min, t_max]


def sample_ndc_pts_on_rays(rays_o, rays_d, xyz_min, xyz_max, N_samples):
    rays_o, rays_d, xyz_min, xyz_max = map(_f, (rays_o, rays_d, xyz_min, xyz_max))
    n = rays_o.shape[0]
    pts = np.empty((n, N_samples, 3), np.float32)
    mask = np.empty((n, N_samples), np.uint8)
    lib().oracle_sample_ndc_pts_on_rays(_p(rays_o), _p(rays_d), _p(xyz_min), _p(xyz_max), N_samples, n,
                                        _p(pts), _p(mask))
    return [pts, mask.astype(bool)]


def sample_bg_pts_on_rays(rays_o, rays_d, t_max, bg_preserve, N_samples):
    rays_o, rays_d, t_max = map(_f, (rays_o, rays_d, t_max))
    n = rays_o.shape[0]
    pts = np.empty((n, N_samples, 3), np.float32)
    lib().oracle_sample_bg_pts_on_rays(_p(rays_o), _p(rays_d), _p(t_max), C.c_float(bg_preserve),
                                       N_samples, n, _p(pts))
    return pts


def maskcache_lookup(world, xyz, scale, shift):
    world = np.ascontiguousarray(world, dtype=np.uint8)
    xyz, scale, shift = map(_f, (xyz, scale, shift))
    out = np.empty(xyz.shape[0], np.uint8)
    lib().oracle_maskcache_lookup(_p(world), _p(xyz), _p(scale), _p(shift), world.shape[0],
                                  world.shape[1], world.shape[2], C.c_int64(xyz.shape[0]), _p(out))
    return out.astype(bool)


def raw2alpha(density, shift, interval):
    density = _f(density)
    exp_d = np.empty_like(density)
    alpha = np.empty_like(density)
    iv = None
    if isinstance(interval, np.ndarray):
        iv = _f(interval)
        interval = 0.0
    lib().oracle_raw2alpha(_p(density), C.c_float(shift), C.c_float(interval), _p(iv),
                           C.c_int64(density.shape[0]), _p(exp_d), _p(alpha))
    return [exp_d, alpha]


def raw2alpha_backward(exp_d, grad_back, interval):
    exp_d, grad_back = map(_f, (exp_d, grad_back))
    grad = np.empty_like(exp_d)
    iv = None
    if isinstance(interval, np.ndarray):
        iv = _f(interval)
        interval = 0.0
    lib().oracle_raw2alpha_backward(_p(exp_d), _p(grad_back), C.c_float(interval), _p(iv),
                                    C.c_int64(exp_d.shape[0]), _p(grad))
    return grad


def alpha2weight(alpha, ray_id, n_rays):
    alpha = _f(alpha)
    ray_id = np.ascontiguousarray(ray_id, dtype=np.int64)
    n = alpha.shape[0]
    weight = np.empty(n, np.float32)
    T = np.empty(n, np.float32)
    last = np.empty(n_rays, np.float32)
    i_start = np.empty(n_rays, np.int64)
    i_end = np.empty(n_rays, np.int64)
    lib().oracle_alpha2weight(_p(alpha), _p(ray_id), C.c_int64(n), n_rays, _p(weight), _p(T), _p(last),
                              _p(i_start), _p(i_end))
    return [weight, T, last, i_start, i_end]


def alpha2weight_backward(alpha, weight, T, alphainv_last, i_start, i_end, n_rays, grad_weights,
                          grad_last):
    alpha, weight, T, alphainv_last, grad_weights, grad_last = map(
        _f, (alpha, weight, T, alphainv_last, grad_weights, grad_last))
    i_start = np.ascontiguousarray(i_start, dtype=np.int64)
    i_end = np.ascontiguousarray(i_end, dtype=np.int64)
    grad = np.empty_like(alpha)
    lib().oracle_alpha2weight_backward(_p(alpha), _p(weight), _p(T), _p(alphainv_last), _p(i_start),
                                       _p(i_end), n_rays, C.c_int64(alpha.shape[0]), _p(grad_weights),
                                       _p(grad_last), _p(grad))
    return grad


def _adam(param, grad, exp_avg, exp_avg_sq, perlr, step, beta1, beta2, lr, eps, mode):
    for a in (param, exp_avg, exp_avg_sq):
        assert a.dtype == np.float32 and a.flags.c_contiguous, "in-place op needs float32 C arrays"
    grad = _f(grad)
    perlr = None if perlr is None else _f(perlr)
    lib().oracle_adam_upd(_p(param), _p(grad), _p(exp_avg), _p(exp_avg_sq), _p(perlr),
                          C.c_int64(param.size), int(step), C.c_float(beta1), C.c_float(beta2),
                          C.c_float(lr), C.c_float(eps), mode)


def adam_upd(param, grad, exp_avg, exp_avg_sq, step, beta1, beta2, lr, eps):
    _adam(param, grad, exp_avg, exp_avg_sq, None, step, beta1, beta2, lr, eps, 0)


def masked_adam_upd(param, grad, exp_avg, exp_avg_sq, step, beta1, beta2, lr, eps):
    _adam(param, grad, exp_avg, exp_avg_sq, None, step, beta1, beta2, lr, eps, 1)


def adam_upd_with_perlr(param, grad, exp_avg, exp_avg_sq, perlr, step, beta1, beta2, lr, eps):
    _adam(param, grad, exp_avg, exp_avg_sq, perlr, step, beta1, beta2, lr, eps, 2)


def total_variation_add_grad(param, grad, wx, wy, wz, dense_mode, mask=None):
    param = _f(param)
    assert grad.dtype == np.float32 and grad.flags.c_contiguous
    assert param.ndim == 5 and param.shape[0] == 1
    m = None if mask is None else _f(mask)
    _, Cn, X, Y, Z = param.shape
    lib().oracle_total_variation_add_grad(_p(param), _p(grad), _p(m), C.c_float(wx), C.c_float(wy),
                                          C.c_float(wz), Cn, X, Y, Z, int(bool(dense_mode)))


def total_variation_add_grad_new(param, grad, mask, wx, wy, wz, dense_mode):
    total_variation_add_grad(param, grad, wx, wy, wz, dense_mode, mask=mask)
