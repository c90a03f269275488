"""GPU accuracy probe: MLP weight-gradient GEMMs (TN) on the real tensors of a fused step, tensor-core
(3xTF32) vs SIMT fp32 vs fp64."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fgs_nerf_b200 import engine, scene
dev = torch.device("cuda:0")
sc = scene.make_scene(G=96)
ro, rd, vd, tgt = [t.to(dev) for t in scene.make_rays(N=4096)]
eng = engine.FineStepEngine(sc, 4096, dev)
eng.forward(ro, rd, vd, 5000); eng.zero_grad(); eng.loss_backward(tgt); torch.cuda.synchronize()
M2 = eng.counters()["M2"]
rows = eng.ws["counters"][1:2]
def rel(a, b): return float((a.double() - b).abs().max() / b.abs().max())
pairs = {"gW0 = dB^T X0": (eng.ws["dB"], eng.ws["X0"]), "gW1 = dA^T H0": (eng.ws["dA"], eng.ws["H"][0])}
for name, (dz, x) in pairs.items():
    want = dz[:M2].double().T @ x[:M2].double()
    absum = dz[:M2].double().abs().T @ x[:M2].double().abs()
    for tc in (False, True):
        out = torch.zeros(dz.shape[1], x.shape[1], device=dev)
        engine.gemm(2, dz, x, out, rows_dev=rows, tc=tc)
        torch.cuda.synchronize()
        print(f"{name}: tc={tc} rel-to-max {rel(out, want):.3e}  max|err|/max(sum|p|) {float((out.double()-want).abs().max()/absum.max()):.3e}  cancellation max|want|/max sum|p| {float(want.abs().max()/absum.max()):.3e}")
    t32 = (dz[:M2].T @ x[:M2])
    print(f"{name}: torch fp32 matmul rel-to-max {rel(t32, want):.3e}")
