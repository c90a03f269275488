"""GPU parity of the Python-class drop-ins (grid.DenseGrid / MaskGrid, adam.MaskedAdam,
nerf.Alphas2Weights / MaskCache) against the reference's classes restated with library ops
(F.grid_sample) and the reference's own compiled kernels (oracle/_ref)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import ref_ext  # noqa: E402

pytestmark = pytest.mark.gpu


def test_dense_grid_forward_backward_and_tv(cuda):
    from fgs_nerf_b200 import grid as G
    g = torch.Generator().manual_seed(3)
    mn, mx = torch.tensor([-1.0, -1.0, -1.0]), torch.tensor([1.0, 1.0, 1.0])
    for C in (1, 12):
        dg = G.create_grid('DenseGrid', channels=C, world_size=torch.tensor([20, 18, 22]), xyz_min=mn, xyz_max=mx).to(cuda)
        dg.grid.data = torch.randn(1, C, 20, 18, 22, generator=g).to(cuda)
        xyz = (torch.rand(4000, 3, generator=g) * 2.2 - 1.1).to(cuda)
        out = dg(xyz)
        ind = ((xyz - mn.to(cuda)) / (mx.to(cuda) - mn.to(cuda))).flip((-1,)) * 2 - 1
        ref_grid = dg.grid.detach().clone().requires_grad_(True)
        want = F.grid_sample(ref_grid, ind.reshape(1, 1, 1, -1, 3), mode='bilinear', align_corners=True)
        want = want.reshape(C, -1).T.reshape(4000, C)
        if C == 1:
            want = want.squeeze(-1)
        assert torch.equal(out, want)                       # forward bit-exact vs ATen
        go = torch.randn(out.shape, generator=g).to(cuda)
        out.backward(go)
        want.backward(go)
        assert float((dg.grid.grad - ref_grid.grad).abs().max()) <= 1e-5 * float(ref_grid.grad.abs().max())
        # TV add-grad through the class method == the reference's kernel on the same tensors
        if ref_ext.available():
            tv = ref_ext.load("total_variation_cuda")
            gref = dg.grid.grad.clone()
            tv.total_variation_add_grad(dg.grid.detach(), gref, 0.3, 0.3, 0.3, True)
            dg.total_variation_add_grad(0.3, 0.3, 0.3, True)
            assert torch.equal(dg.grid.grad, gref)
        assert tuple(dg.state_dict()['grid'].shape) == (1, C, 20, 18, 22)
    want = F.interpolate(dg.grid.data, size=(30, 30, 30), mode='trilinear', align_corners=True)   # model/grid.py:106
    dg.scale_volume_grid(torch.tensor([30, 30, 30]))
    assert tuple(dg.grid.shape) == (1, 12, 30, 30, 30) and isinstance(dg.grid, torch.nn.Parameter)
    assert float((dg.grid.data - want).abs().max()) <= 2e-6 * float(want.abs().max())


def test_mask_grid_and_mask_cache(cuda):
    from fgs_nerf_b200 import grid as G, nerf as Nf
    g = torch.Generator().manual_seed(4)
    mask = torch.rand(24, 20, 28, generator=g) > 0.5
    mg = G.MaskGrid(mask=mask, xyz_min=[-1.0, -1.0, -1.0], xyz_max=[1.0, 1.0, 1.0]).to(cuda)
    xyz = (torch.rand(5000, 3, generator=g) * 2.4 - 1.2).to(cuda)
    got = mg(xyz)
    ijk = torch.round(xyz * mg.xyz2ijk_scale + mg.xyz2ijk_shift).long()
    inb = ((ijk >= 0) & (ijk < torch.tensor(list(mask.shape), device=cuda))).all(-1)
    want = torch.zeros(5000, dtype=torch.bool, device=cuda)
    want[inb] = mask.to(cuda)[ijk[inb, 0], ijk[inb, 1], ijk[inb, 2]]
    agree = float((got == want).float().mean())
    assert agree > 0.999, agree          # torch evaluates xyz*scale+shift unfused; ties at .5 may differ
    sdf_mask = (torch.rand(1, 1, 30, 30, 30, generator=g) * 2e-3).to(cuda)
    mc = Nf.MaskCache(mask_cache_thres=1e-3, sdf_mask=sdf_mask, xyz_min=[-1.2] * 3, xyz_max=[1.2] * 3).to(cuda)
    ind = ((xyz - mc.xyz_min) / (mc.xyz_max - mc.xyz_min)).flip((-1,)) * 2 - 1
    ref = F.grid_sample(sdf_mask, ind.reshape(1, 1, 1, -1, 3), align_corners=True).reshape(-1) >= 1e-3
    assert torch.equal(mc(xyz), ref)


def test_alphas2weights_autograd(cuda):
    from fgs_nerf_b200 import nerf as Nf
    if not ref_ext.available():
        pytest.skip("oracle/_ref not built")
    ru = ref_ext.load("render_utils_cuda")
    g = torch.Generator().manual_seed(5)
    lens = torch.randint(0, 40, (300,), generator=g)
    ray_id = torch.repeat_interleave(torch.arange(300), lens).to(cuda)
    alpha = (torch.rand(ray_id.numel(), generator=g) ** 3).to(cuda).requires_grad_(True)
    w, last = Nf.Alphas2Weights.apply(alpha, ray_id, 300)
    gw, gl = torch.randn(w.shape, generator=g).to(cuda), torch.randn(300, generator=g).to(cuda)
    (w * gw).sum().backward(retain_graph=True)
    g1 = alpha.grad.clone()
    alpha.grad = None
    ((w * gw).sum() + (last * gl).sum()).backward()
    rw, rT, rlast, ris, rie = ru.alpha2weight(alpha.detach(), ray_id, 300)
    assert torch.equal(w, rw) and torch.equal(last, rlast)
    rg = ru.alpha2weight_backward(alpha.detach(), rw, rT, rlast, ris, rie, 300, gw, gl)
    assert torch.equal(alpha.grad, rg)
    assert not torch.equal(g1, alpha.grad)


def test_masked_adam_matches_reference_kernels(cuda):
    from fgs_nerf_b200.adam import MaskedAdam
    if not ref_ext.available():
        pytest.skip("oracle/_ref not built")
    ad = ref_ext.load("adam_upd_cuda")
    g = torch.Generator().manual_seed(6)
    k0 = torch.randn(1, 12, 10, 11, 12, generator=g).to(cuda).requires_grad_(True)
    sdf = torch.randn(1, 1, 10, 11, 12, generator=g).to(cuda).requires_grad_(True)
    ws = [torch.randn(256, 106, generator=g).to(cuda).requires_grad_(True), torch.randn(256, generator=g).to(cuda).requires_grad_(True)]
    opt = MaskedAdam([dict(params=[k0], lr=0.1, skip_zero_grad=True), dict(params=[sdf], lr=0.005, skip_zero_grad=False),
                      dict(params=ws, lr=1e-3, skip_zero_grad=False)], betas=(0.9, 0.99))
    ref = {id(p): (p.detach().clone(), torch.zeros_like(p), torch.zeros_like(p)) for p in [k0, sdf] + ws}
    for step in range(1, 4):
        for p in [k0, sdf] + ws:
            p.grad = torch.randn(p.shape, generator=g).to(cuda)
        k0.grad[torch.rand(k0.shape, generator=g).to(cuda) < 0.8] = 0.0
        opt.step()
        for p, lr, masked in ((k0, 0.1, True), (sdf, 0.005, False), (ws[0], 1e-3, False), (ws[1], 1e-3, False)):
            rp, rm, rv = ref[id(p)]
            (ad.masked_adam_upd if masked else ad.adam_upd)(rp, p.grad, rm, rv, step, 0.9, 0.99, lr, 1e-8)
            assert torch.equal(p.detach(), rp), f"step {step}"
            assert torch.equal(opt.state[p]['exp_avg'], rm) and torch.equal(opt.state[p]['exp_avg_sq'], rv)
