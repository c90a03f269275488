"""GPU parity of the drop-in native ops: our sm_100a kernels (through the C-ABI, via the
reference-named Python modules) vs (a) the CPU oracle and (b) the reference's own kernels
(oracle/_ref/*.so) on the same seeded inputs.  ids / masks / integer outputs and every fp32
output whose arithmetic is pure IEEE add/mul/fma/div/sqrt must be BIT-EXACT."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import native_cases as nc  # noqa: E402
from oracle import native_ref as orc  # noqa: E402
from oracle import ref_ext  # noqa: E402

pytestmark = pytest.mark.gpu


def _mods():
    from fgs_nerf_b200 import adam_upd_cuda, render_utils_cuda, total_variation_cuda
    return render_utils_cuda, adam_upd_cuda, total_variation_cuda


def _cu(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def _eq(a, b, what):
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    b = b.detach().cpu().numpy() if torch.is_tensor(b) else np.asarray(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    if a.dtype.kind == "f":
        same = (a.view(np.uint32) == b.view(np.uint32)) | ((a == 0) & (b == 0))
    else:
        same = a == b
    assert same.all(), f"{what}: {int((~same).sum())} of {a.size} elements differ (bit-exact required)"


def _refs(cuda):
    """oracle functions + (optionally) the reference's compiled modules."""
    out = [("oracle", None)]
    if ref_ext.available():
        out.append(("reference", (ref_ext.load("render_utils_cuda"), ref_ext.load("adam_upd_cuda"),
                                  ref_ext.load("total_variation_cuda"))))
    return out


def test_library_is_native(cuda):
    from fgs_nerf_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH)
    assert _lib.lib.fgs_abi_version() == 1
    assert any("libfgs_b200.so" in l for l in open("/proc/self/maps"))


@pytest.mark.parametrize("n", [384, 1, 5])
def test_sample_pts_on_rays(cuda, n):
    ru, _, _ = _mods()
    c = nc.sampling_case(n=max(n, 5))
    for k in ("rays_o", "rays_d"):
        c[k] = np.ascontiguousarray(c[k][:n])
    args_np = (c["rays_o"], c["rays_d"], c["xyz_min"], c["xyz_max"], float(c["near"]), float(c["far"]), float(c["stepdist"]))
    args_cu = tuple(_cu(a, cuda) if isinstance(a, np.ndarray) else a for a in args_np)
    ours = ru.sample_pts_on_rays(*args_cu)
    names = ["ray_pts", "mask_outbbox", "ray_id", "step_id", "N_steps", "t_min", "t_max"]
    want = orc.sample_pts_on_rays(*args_np)
    for nme, a, b in zip(names, ours, want):
        _eq(a, b, f"sample_pts_on_rays.{nme} vs oracle")
    if ref_ext.available():
        ref = ref_ext.load("render_utils_cuda").sample_pts_on_rays(*args_cu)
        for nme, a, b in zip(names, ours, ref):
            _eq(a, b, f"sample_pts_on_rays.{nme} vs reference")
    # the split ops agree with the fused one
    t_min, t_max = ru.infer_t_minmax(*args_cu[:6])
    _eq(t_min, ours[5], "infer_t_minmax.t_min")
    _eq(t_max, ours[6], "infer_t_minmax.t_max")
    _eq(ru.infer_n_samples(args_cu[1], t_min, t_max, args_np[6]), ours[4], "infer_n_samples")
    s, d = ru.infer_ray_start_dir(args_cu[0], args_cu[1], t_min)
    so, do = orc.infer_ray_start_dir(c["rays_o"], c["rays_d"], want[5])
    _eq(s, so, "rays_start")
    _eq(d, do, "rays_dir")


def test_sample_pts_scene_rays_full_size(cuda):
    """BASELINE config (2) batch: 8192 blender rays at 160^3 -- ours vs the reference kernels,
    plus size-independent invariants (segments contiguous, step ids 0..n-1, counts add up)."""
    ru, _, _ = _mods()
    from fgs_nerf_b200 import scene
    ro, rd, _, _ = scene.make_rays(N=8192)
    G = 160
    voxel = float((torch.tensor(8.0) / G ** 3).pow(1 / 3))
    mn, mx = torch.tensor([-1., -1, -1]).to(cuda), torch.tensor([1., 1, 1]).to(cuda)
    args = (ro.to(cuda), rd.to(cuda), mn, mx, 2.0, 1e9, 0.5 * voxel)
    ours = ru.sample_pts_on_rays(*args)
    ray_id, step_id, n_steps = ours[2], ours[3], ours[4]
    assert int(n_steps.sum()) == ray_id.numel()
    assert (ray_id[1:] >= ray_id[:-1]).all()
    assert torch.equal(torch.bincount(ray_id, minlength=8192), n_steps)
    first = torch.ones_like(ray_id, dtype=torch.bool)
    first[1:] = ray_id[1:] != ray_id[:-1]
    assert (step_id[first] == 0).all()
    assert ((step_id[1:] - step_id[:-1])[~first[1:]] == 1).all()
    if ref_ext.available():
        ref = ref_ext.load("render_utils_cuda").sample_pts_on_rays(*args)
        for nme, a, b in zip(["ray_pts", "mask_outbbox", "ray_id", "step_id", "N_steps", "t_min", "t_max"], ours, ref):
            _eq(a, b, f"full-size sample_pts_on_rays.{nme} vs reference")


def test_sample_ndc_bg(cuda):
    ru, _, _ = _mods()
    c = nc.sampling_case()
    ro, rd = c["rays_o"][:64].copy(), c["rays_d"][:64].copy()
    p, m = ru.sample_ndc_pts_on_rays(_cu(ro, cuda), _cu(rd, cuda), _cu(c["xyz_min"], cuda), _cu(c["xyz_max"], cuda), 17)
    po, mo = orc.sample_ndc_pts_on_rays(ro, rd, c["xyz_min"], c["xyz_max"], 17)
    _eq(p, po, "ndc pts")
    _eq(m, mo, "ndc mask")
    t_max = orc.infer_t_minmax(ro, rd, c["xyz_min"], c["xyz_max"], float(c["near"]), float(c["far"]))[1]
    b = ru.sample_bg_pts_on_rays(_cu(ro, cuda), _cu(rd, cuda), _cu(t_max, cuda), 0.5, 16)
    _eq(b, orc.sample_bg_pts_on_rays(ro, rd, t_max, 0.5, 16), "bg pts")
    if ref_ext.available():
        r = ref_ext.load("render_utils_cuda")
        _eq(b, r.sample_bg_pts_on_rays(_cu(ro, cuda), _cu(rd, cuda), _cu(t_max, cuda), 0.5, 16), "bg pts vs reference")
        pr, mr = r.sample_ndc_pts_on_rays(_cu(ro, cuda), _cu(rd, cuda), _cu(c["xyz_min"], cuda), _cu(c["xyz_max"], cuda), 17)
        _eq(p, pr, "ndc pts vs reference")
        _eq(m, mr, "ndc mask vs reference")


def test_maskcache_lookup(cuda):
    ru, _, _ = _mods()
    c = nc.maskcache_case()
    got = ru.maskcache_lookup(_cu(c["world"], cuda), _cu(c["xyz"], cuda), _cu(c["scale"], cuda), _cu(c["shift"], cuda))
    assert got.dtype == torch.bool
    _eq(got, orc.maskcache_lookup(c["world"], c["xyz"], c["scale"], c["shift"]), "maskcache vs oracle")
    if ref_ext.available():
        ref = ref_ext.load("render_utils_cuda").maskcache_lookup(
            _cu(c["world"], cuda), _cu(c["xyz"], cuda), _cu(c["scale"], cuda), _cu(c["shift"], cuda))
        _eq(got, ref, "maskcache vs reference")
    empty = ru.maskcache_lookup(_cu(c["world"], cuda), torch.empty((0, 3), device=cuda), _cu(c["scale"], cuda), _cu(c["shift"], cuda))
    assert empty.shape == (0,)


def test_raw2alpha(cuda):
    ru, _, _ = _mods()
    c = nc.raw2alpha_case()
    d, gb, iv = _cu(c["density"], cuda), _cu(c["grad_back"], cuda), _cu(c["interval_pts"], cuda)
    e, a = ru.raw2alpha(d, float(c["shift"]), float(c["interval"]))
    g = ru.raw2alpha_backward(e, gb, float(c["interval"]))
    e2, a2 = ru.raw2alpha_nonuni(d, float(c["shift"]), iv)
    g2 = ru.raw2alpha_nonuni_backward(e2, gb, iv)
    eo, ao = orc.raw2alpha(c["density"], float(c["shift"]), float(c["interval"]))
    # libm expf/powf vs CUDA's: a few ulp
    np.testing.assert_allclose(a.cpu().numpy(), ao, rtol=2e-6, atol=2e-7)
    np.testing.assert_allclose(g.cpu().numpy(), orc.raw2alpha_backward(eo, c["grad_back"], float(c["interval"])), rtol=1e-5, atol=1e-12)
    if ref_ext.available():
        r = ref_ext.load("render_utils_cuda")
        er, ar = r.raw2alpha(d, float(c["shift"]), float(c["interval"]))
        _eq(e, er, "raw2alpha exp vs reference")
        _eq(a, ar, "raw2alpha alpha vs reference")
        _eq(g, r.raw2alpha_backward(er, gb, float(c["interval"])), "raw2alpha_backward vs reference")
        er2, ar2 = r.raw2alpha_nonuni(d, float(c["shift"]), iv)
        _eq(a2, ar2, "raw2alpha_nonuni vs reference")
        _eq(g2, r.raw2alpha_nonuni_backward(er2, gb, iv), "raw2alpha_nonuni_backward vs reference")


@pytest.mark.parametrize("seed,n_rays,max_len", [(1, 97, 70), (11, 400, 300), (12, 3, 5)])
def test_alpha2weight(cuda, seed, n_rays, max_len):
    ru, _, _ = _mods()
    c = nc.alpha_case(n_rays=n_rays, seed=seed, max_len=max_len)
    al, rid = _cu(c["alpha"], cuda), _cu(c["ray_id"], cuda)
    ours = ru.alpha2weight(al, rid, c["n_rays"])
    want = orc.alpha2weight(c["alpha"], c["ray_id"], c["n_rays"])
    names = ["weight", "T", "alphainv_last", "i_start", "i_end"]
    for nme, a, b in zip(names, ours, want):
        _eq(a, b, f"alpha2weight.{nme} vs oracle")
    gw, gl = _cu(c["grad_weights"], cuda), _cu(c["grad_last"], cuda)
    g = ru.alpha2weight_backward(al, *ours, c["n_rays"], gw, gl)
    go = orc.alpha2weight_backward(c["alpha"], *want, c["n_rays"], c["grad_weights"], c["grad_last"])
    _eq(g, go, "alpha2weight_backward vs oracle")
    if ref_ext.available():
        r = ref_ext.load("render_utils_cuda")
        ref = r.alpha2weight(al, rid, c["n_rays"])
        for nme, a, b in zip(names, ours, ref):
            _eq(a, b, f"alpha2weight.{nme} vs reference")
        _eq(g, r.alpha2weight_backward(al, *ref, c["n_rays"], gw, gl), "alpha2weight_backward vs reference")


def test_alpha2weight_empty(cuda):
    ru, _, _ = _mods()
    out = ru.alpha2weight(torch.empty(0, device=cuda), torch.empty(0, dtype=torch.int64, device=cuda), 9)
    assert out[0].numel() == 0 and out[1].numel() == 0
    assert torch.equal(out[2], torch.ones(9, device=cuda))
    assert int(out[3].abs().sum()) == 0 and int(out[4].abs().sum()) == 0


@pytest.mark.parametrize("n", [5003, 4096, 3])
def test_adam(cuda, n):
    _, ad, _ = _mods()
    c = nc.adam_case(n=n)
    for mode in ("dense", "masked", "perlr"):
        p, m, v = (_cu(c[k].copy(), cuda) for k in ("param", "exp_avg", "exp_avg_sq"))
        po, mo, vo = (c[k].copy() for k in ("param", "exp_avg", "exp_avg_sq"))
        if ref_ext.available():
            pr, mr, vr = (_cu(c[k].copy(), cuda) for k in ("param", "exp_avg", "exp_avg_sq"))
            r = ref_ext.load("adam_upd_cuda")
        for s, g in enumerate(c["grads"]):
            step = c["step0"] + s
            hp = (step, c["beta1"], c["beta2"], c["lr"], c["eps"])
            if mode == "dense":
                ad.adam_upd(p, _cu(g, cuda), m, v, *hp); orc.adam_upd(po, g, mo, vo, *hp)
                if ref_ext.available(): r.adam_upd(pr, _cu(g, cuda), mr, vr, *hp)
            elif mode == "masked":
                ad.masked_adam_upd(p, _cu(g, cuda), m, v, *hp); orc.masked_adam_upd(po, g, mo, vo, *hp)
                if ref_ext.available(): r.masked_adam_upd(pr, _cu(g, cuda), mr, vr, *hp)
            else:
                ad.adam_upd_with_perlr(p, _cu(g, cuda), m, v, _cu(c["perlr"], cuda), *hp)
                orc.adam_upd_with_perlr(po, g, mo, vo, c["perlr"], *hp)
                if ref_ext.available(): r.adam_upd_with_perlr(pr, _cu(g, cuda), mr, vr, _cu(c["perlr"], cuda), *hp)
        for nme, a, b in (("param", p, po), ("exp_avg", m, mo), ("exp_avg_sq", v, vo)):
            _eq(a, b, f"adam[{mode}].{nme} vs oracle")
        if ref_ext.available():
            for nme, a, b in (("param", p, pr), ("exp_avg", m, mr), ("exp_avg_sq", v, vr)):
                _eq(a, b, f"adam[{mode}].{nme} vs reference")


def test_adam_unaligned_view_and_multi(cuda):
    _, ad, _ = _mods()
    c = nc.adam_case(n=1001)
    hp = (3, 0.9, 0.99, 1e-3, 1e-8)
    base = [_cu(np.concatenate([[0], c[k]]).astype(np.float32), cuda) for k in ("param", "exp_avg", "exp_avg_sq")]
    p, m, v = (b[1:] for b in base)           # 4-byte-offset views: scalar path
    g = _cu(np.concatenate([[0], c["grads"][0]]).astype(np.float32), cuda)[1:]
    ad.adam_upd(p, g, m, v, *hp)
    po, mo, vo = (c[k].copy() for k in ("param", "exp_avg", "exp_avg_sq"))
    orc.adam_upd(po, c["grads"][0], mo, vo, *hp)
    _eq(p, po, "adam view param")
    # multi-tensor == per-tensor
    rs = np.random.RandomState(0)
    shapes = [(256, 106), (256,), (256, 256), (3, 256), (3,)]
    mk = lambda: [_cu(rs.normal(size=s).astype(np.float32), cuda) for s in shapes]
    P, G, M = mk(), mk(), mk()
    V = [x.abs() for x in mk()]
    P2, M2, V2 = [x.clone() for x in P], [x.clone() for x in M], [x.clone() for x in V]
    ad.adam_upd_multi(P, G, M, V, [5] * 5, [1e-3] * 5, 0.9, 0.99, 1e-8)
    for a, b, cc, d in zip(P2, G, M2, V2):
        ad.adam_upd(a, b, cc, d, 5, 0.9, 0.99, 1e-3, 1e-8)
    for a, b in zip(P + M + V, P2 + M2 + V2):
        _eq(a, b, "adam multi vs single")


@pytest.mark.parametrize("C,shape", [(3, (9, 6, 12)), (1, (8, 8, 16)), (1, (5, 7, 9)), (12, (6, 5, 8))])
def test_total_variation(cuda, C, shape):
    _, _, tv = _mods()
    c = nc.tv_case(C=C, shape=shape)
    P = _cu(c["param"], cuda)
    for dense in (True, False):
        g = _cu(c["grad"].copy(), cuda)
        tv.total_variation_add_grad(P, g, c["wx"], c["wy"], c["wz"], dense)
        go = c["grad"].copy()
        orc.total_variation_add_grad(c["param"], go, c["wx"], c["wy"], c["wz"], dense)
        _eq(g, go, f"tv dense={dense} vs oracle")
        gm = _cu(c["grad"].copy(), cuda)
        tv.total_variation_add_grad_new(P, gm, _cu(c["mask"], cuda), c["wx"], c["wy"], c["wz"], dense)
        gmo = c["grad"].copy()
        orc.total_variation_add_grad_new(c["param"], gmo, c["mask"], c["wx"], c["wy"], c["wz"], dense)
        _eq(gm, gmo, f"tv masked dense={dense} vs oracle")
        if ref_ext.available():
            r = ref_ext.load("total_variation_cuda")
            gr = _cu(c["grad"].copy(), cuda)
            r.total_variation_add_grad(P, gr, c["wx"], c["wy"], c["wz"], dense)
            _eq(g, gr, f"tv dense={dense} vs reference")
            gr = _cu(c["grad"].copy(), cuda)
            r.total_variation_add_grad_new(P, gr, _cu(c["mask"], cuda), c["wx"], c["wy"], c["wz"], dense)
            _eq(gm, gr, f"tv masked dense={dense} vs reference")
        # channels_last_3d storage of the same logical tensor gives the same logical result
        Pcl = P.contiguous(memory_format=torch.channels_last_3d)
        gcl = _cu(c["grad"].copy(), cuda).contiguous(memory_format=torch.channels_last_3d)
        if C > 1:
            assert not Pcl.is_contiguous()
        tv.total_variation_add_grad(Pcl, gcl, c["wx"], c["wy"], c["wz"], dense)
        _eq(gcl.contiguous(), go, f"tv channels_last dense={dense}")


def test_error_behaviour(cuda):
    """CHECK_INPUT semantics of render_utils.cpp:46-48 -> RuntimeError."""
    ru, ad, _ = _mods()
    c = nc.sampling_case(n=8)
    ro, rd = _cu(c["rays_o"], cuda), _cu(c["rays_d"], cuda)
    mn, mx = _cu(c["xyz_min"], cuda), _cu(c["xyz_max"], cuda)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        ru.sample_pts_on_rays(ro.cpu(), rd, mn, mx, 0.2, 1e9, 0.01)
    with pytest.raises(RuntimeError, match="must be contiguous"):
        ru.sample_pts_on_rays(ro.t().contiguous().t(), rd, mn, mx, 0.2, 1e9, 0.01)
    with pytest.raises(RuntimeError):
        ad.adam_upd(ro.double(), ro, ro, ro, 1, 0.9, 0.99, 1e-3, 1e-8)


def test_voxel_list_and_copy():
    """fgs_voxel_list (ascending ids of the flagged voxels, flags cleared, two launches over all SMs) and
    fgs_voxel_copy (gather / scatter of 12 channels per listed voxel) against torch."""
    import torch
    from fgs_nerf_b200._lib import lib, check, ptr, stream
    dev = torch.device("cuda:0")
    for n, p in ((48 ** 3, 0.02), (16384 * 3 + 777, 0.3), (1000, 0.0), (160 ** 3, 0.03)):
        g = torch.Generator(device="cpu").manual_seed(n)
        flags = (torch.rand(n, generator=g) < p).to(torch.uint8).to(dev)
        want = flags.nonzero()[:, 0].to(torch.int32)
        lst = torch.zeros(int(lib.fgs_voxel_list_ints(n)), dtype=torch.int32, device=dev)
        check(lib.fgs_voxel_list(ptr(flags), n, ptr(lst), stream()))
        torch.cuda.synchronize()
        cnt = int(lst[0])
        assert cnt == want.numel()
        assert torch.equal(lst[1:1 + cnt], want)
        assert int(flags.sum()) == 0, "flags must be cleared"
        if cnt:
            grid = torch.randn(n, 12, generator=g).to(dev)
            compact = torch.zeros(cnt, 12, device=dev)
            check(lib.fgs_voxel_copy(ptr(grid), ptr(compact), ptr(lst), cnt, 12, 1, stream()))
            torch.cuda.synchronize()
            assert torch.equal(compact, grid[want.long()])
            grid2 = torch.zeros_like(grid)
            check(lib.fgs_voxel_copy(ptr(grid2), ptr(compact), ptr(lst), cnt, 12, 0, stream()))
            torch.cuda.synchronize()
            ref = torch.zeros_like(grid)
            ref[want.long()] = grid[want.long()]
            assert torch.equal(grid2, ref)


def test_voxel_bit_masks_union_list():
    """fgs_voxel_bits_pack + fgs_voxel_bits_union_list: byte flags -> bit masks, OR over `world` masks, touched-brick
    flags of the union, ascending voxel list -- against torch."""
    import torch
    from fgs_nerf_b200._lib import lib, check, ptr, stream
    dev = torch.device("cuda:0")
    for (X, Y, Z), world in (((48, 48, 48), 2), ((16, 20, 32), 3), ((160, 160, 160), 8)):
        V = X * Y * Z
        g = torch.Generator(device="cpu").manual_seed(V + world)
        masks = [(torch.rand(V, generator=g) < 0.01) for _ in range(world)]
        bits = torch.zeros(world, V // 32, dtype=torch.int32, device=dev)
        for r, m in enumerate(masks):
            flags = m.to(torch.uint8).to(dev)
            check(lib.fgs_voxel_bits_pack(ptr(flags), ptr(bits[r]), V, stream()))
            torch.cuda.synchronize()
            assert int(flags.sum()) == 0, "byte flags must be cleared"
        union = torch.stack(masks).any(0).to(dev)
        want = union.nonzero()[:, 0].to(torch.int32)
        nb = ((X + 3) // 4, (Y + 3) // 4, (Z + 3) // 4)
        brick_flags = torch.zeros(nb[0] * nb[1] * nb[2], dtype=torch.uint8, device=dev)
        lst = torch.zeros(int(lib.fgs_voxel_list_ints(V)), dtype=torch.int32, device=dev)
        check(lib.fgs_voxel_bits_union_list(ptr(bits), world, X, Y, Z, ptr(brick_flags), ptr(lst), stream()))
        torch.cuda.synchronize()
        cnt = int(lst[0])
        assert cnt == want.numel() and torch.equal(lst[1:1 + cnt], want)
        u3 = union.reshape(X, Y, Z)
        want_bricks = torch.nn.functional.max_pool3d(u3[None, None].float(), 4, 4, ceil_mode=True)[0, 0].reshape(-1) > 0
        assert torch.equal(brick_flags.bool(), want_bricks)
