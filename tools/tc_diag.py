"""GPU diagnostic: tensor-core vs SIMT GEMM backends on the buffers of a fused step (config 1)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fgs_nerf_b200 import engine, scene
dev = torch.device("cuda:0")
G, N = int(sys.argv[1]) if len(sys.argv) > 1 else 96, int(sys.argv[2]) if len(sys.argv) > 2 else 4096
sc = scene.make_scene(G=G)
ro, rd, vd, tgt = [t.to(dev) for t in scene.make_rays(N=N)]
snap = {}
for tc in (False, True):
    engine.set_gemm_backend(tc)
    eng = engine.FineStepEngine(sc, N, dev)
    eng.forward(ro, rd, vd, 5000); eng.zero_grad(); eng.loss_backward(tgt); torch.cuda.synchronize()
    M2 = eng.counters()["M2"]
    d = {}
    for i in range(6): d[f"H{i}"] = eng.ws["H"][i][:M2].clone()
    d["XR"] = eng.ws["XR"][:M2, :256].clone(); d["logit"] = eng.ws["logit"][:M2, :3].clone()
    d["dlogit"] = eng.ws["dlogit"][:M2, :3].clone()
    d["dXR"] = eng.ws["dXR"][:M2, :307].clone(); d["dX0"] = eng.ws["dX0"][:M2, :106].clone()
    d["dA(last)"] = eng.ws["dA"][:M2].clone(); d["dB(last)"] = eng.ws["dB"][:M2].clone()
    for i in range(8): d[f"gW{i}"] = eng.gW[i].clone(); d[f"gB{i}"] = eng.gB[i].clone()
    d["g_k0"] = eng.g_k0.clone(); d["g_sdf"] = eng.g_sdf.clone()
    snap[tc] = d
print("M2 =", M2)
for k in snap[True]:
    a, b = snap[True][k].double(), snap[False][k].double()
    err = (a - b).abs()
    mx = float(b.abs().max())
    flat = int(err.reshape(-1).argmax())
    row = flat // a.shape[-1] if a.dim() == 2 else -1
    nbad = int((err > 1e-4 * mx).sum())
    sign = int(((a > 0) != (b > 0)).sum()) if k.startswith("H") else -1
    print(f"{k:9s} max|err|/max {float(err.max())/mx:.3e}  at row {row}  n(err>1e-4 max) {nbad}  relu sign flips {sign}")
