"""Seeded inputs shared by the golden-vector generator, the CPU oracle tests and the GPU parity
tests of the native ops (reference: model/cuda/render_utils*, adam_upd*, total_variation*)."""
import numpy as np


def sampling_case(n=384, seed=0, G=96):
    """Rays around a [-1,1]^3 box: blender-like hits, grazing rays, misses, axis-parallel rays
    (zero direction components exercise the 1e-6 substitution of render_utils_kernel.cu:23-25)."""
    rs = np.random.RandomState(seed)
    o = rs.normal(size=(n, 3)).astype(np.float32)
    o = o / np.linalg.norm(o, axis=-1, keepdims=True) * rs.uniform(2.5, 4.5, size=(n, 1)).astype(np.float32)
    tgt = rs.uniform(-1.3, 1.3, size=(n, 3)).astype(np.float32)
    d = (tgt - o).astype(np.float32)
    d *= rs.uniform(0.3, 2.0, size=(n, 1)).astype(np.float32)   # un-normalised like pinhole rays
    # axis-parallel / zero components
    d[0] = [0.0, 0.0, -1.0]; o[0] = [0.1, 0.2, 3.0]
    d[1] = [1.0, 0.0, 0.0]; o[1] = [-3.0, 0.5, -0.5]
    d[2] = [0.0, 1.0, 0.0]; o[2] = [5.0, -3.0, 0.0]          # misses the box
    d[3] = [-0.0, 0.7, 0.7]; o[3] = [0.0, -3.0, -3.0]
    o[4] = [0.0, 0.0, 0.0]; d[4] = [0.3, 0.4, 0.5]            # origin inside the box
    xyz_min = np.array([-1, -1, -1], np.float32)
    xyz_max = np.array([1, 1, 1], np.float32)
    voxel = np.float32((8.0 / G ** 3) ** (1 / 3))
    return dict(rays_o=o, rays_d=d.astype(np.float32), xyz_min=xyz_min, xyz_max=xyz_max,
                near=np.float32(0.2), far=np.float32(1e9), stepdist=np.float32(0.5 * voxel))


def alpha_case(n_rays=97, seed=1, max_len=70):
    """Sorted ray_id with empty rays, long opaque runs (early stop T<1e-3) and faint runs."""
    rs = np.random.RandomState(seed)
    lens = rs.randint(0, max_len, size=n_rays)
    lens[rs.rand(n_rays) < 0.15] = 0
    if n_rays > 8:
        lens[5] = 0
        lens[7] = 1
    lens[n_rays - 1] = 0          # trailing empty ray
    ray_id = np.repeat(np.arange(n_rays), lens).astype(np.int64)
    kind = rs.randint(0, 3, size=n_rays)
    alpha = np.empty(ray_id.shape[0], np.float32)
    for r in range(n_rays):
        m = ray_id == r
        k = int(m.sum())
        if kind[r] == 0:
            a = rs.uniform(0.0, 0.05, size=k)
        elif kind[r] == 1:
            a = rs.uniform(0.0, 0.9, size=k)
        else:
            a = rs.beta(0.5, 0.5, size=k)
        alpha[m] = a.astype(np.float32)
    if alpha.size > 3:
        alpha[0] = 1.0   # fully opaque first sample
        alpha[1] = 0.0
    gw = rs.normal(size=alpha.shape[0]).astype(np.float32)
    gl = rs.normal(size=n_rays).astype(np.float32)
    return dict(alpha=alpha, ray_id=ray_id, n_rays=n_rays, grad_weights=gw, grad_last=gl)


def maskcache_case(seed=2, n=3000, shape=(17, 23, 19)):
    rs = np.random.RandomState(seed)
    world = rs.rand(*shape) < 0.4
    xyz_min = np.array([-1.0, -0.5, -2.0], np.float32)
    xyz_max = np.array([1.0, 1.5, 0.5], np.float32)
    scale = ((np.array(shape, np.float32) - 1) / (xyz_max - xyz_min)).astype(np.float32)
    shift = (-xyz_min * scale).astype(np.float32)
    xyz = (rs.uniform(-0.2, 1.2, size=(n, 3)) * (xyz_max - xyz_min) + xyz_min).astype(np.float32)
    return dict(world=world, xyz=xyz, scale=scale, shift=shift)


def adam_case(seed=3, n=5003):
    rs = np.random.RandomState(seed)
    p = rs.normal(size=n).astype(np.float32)
    m = (rs.normal(size=n) * 0.01).astype(np.float32)
    v = (rs.uniform(size=n) * 1e-4).astype(np.float32)
    grads = []
    for _ in range(3):
        g = rs.normal(size=n).astype(np.float32) * 1e-2
        g[rs.rand(n) < 0.5] = 0.0
        grads.append(g)
    perlr = rs.uniform(size=n).astype(np.float32)
    return dict(param=p, exp_avg=m, exp_avg_sq=v, grads=grads, perlr=perlr,
                beta1=0.9, beta2=0.99, lr=0.1, eps=1e-8, step0=7)


def tv_case(seed=4, C=3, shape=(9, 6, 12)):
    rs = np.random.RandomState(seed)
    p = (rs.normal(size=(1, C) + shape) * 1.5).astype(np.float32)   # some |diff| > 1 -> clamp
    g = rs.normal(size=(1, C) + shape).astype(np.float32)
    g[rs.rand(*g.shape) < 0.5] = 0.0
    mask = (rs.rand(*g.shape) < 0.7).astype(np.float32)
    return dict(param=p, grad=g, mask=mask, wx=0.3, wy=0.5, wz=0.7)


def raw2alpha_case(seed=5, n=2000):
    rs = np.random.RandomState(seed)
    d = (rs.normal(size=n) * 4).astype(np.float32)
    d[:3] = [80.0, -80.0, 0.0]
    iv = rs.uniform(0.1, 1.0, size=n).astype(np.float32)
    gb = rs.normal(size=n).astype(np.float32)
    return dict(density=d, shift=np.float32(-4.0), interval=np.float32(0.5), interval_pts=iv, grad_back=gb)
