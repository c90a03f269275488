"""GPU probe: per-role phase cycles of CTA 0 of the fused MLP chain kernels on the bench workload."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fgs_nerf_b200 import engine, scene
from fgs_nerf_b200._lib import lib, check
dev = torch.device("cuda:0")
G = int(sys.argv[1]) if len(sys.argv) > 1 else 160
# (the chain always runs on CTA pairs)
N = 8192
sc = scene.make_scene(G=G)
ro, rd, vd, tgt = [t.to(dev) for t in scene.make_rays(N=N)]
eng = engine.FineStepEngine(sc, N, dev)
eng.set_nonempty_mask()
for i in range(3):
    eng.train_step(ro, rd, vd, tgt, 5000 + i)
torch.cuda.synchronize()
names = ["consumer: wait r_full", "consumer: epilogue (+store wait)", "producer: wait c_done (dependency)",
         "producer: wait stage free", "producer: split/store/load issue", "issuer: wait a_full", "issuer: wait b_full",
         "issuer: wait r_empty", "issuer: wait free (B prefetch)", "issuer: total"]
dbg = torch.zeros(32, dtype=torch.int64, device=dev)
def show(tag, ms, off=0):
    d = dbg.cpu().tolist()[off:off + 10]; tot = max(d[9], 1)
    print(f"{tag}: {ms*1e3:.1f} us, M2 = {eng.counters()['M2']}")
    for n, v in zip(names, d):
        print(f"  {n:36s} {v:9d} cyc  {100*v/tot:5.1f}%   {v/1965:.2f} us")
    if off == 0:
        e = dbg.cpu().tolist()[20:25]
        for n, v in zip(["epilogue: bias fill + barrier", "epilogue: TMEM loads", "epilogue: bits + arithmetic",
                         "epilogue: q loop total (from r_full)", "epilogue: fence + arrivals"], e):
            print(f"    {n:36s} {v:9d} cyc  {v/1965:.2f} us")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
check(lib.fgs_debug_gemm_phases(dbg.data_ptr()))
e0.record(); eng.forward(ro, rd, vd, 5003, _skip_composite=True); e1.record(); torch.cuda.synchronize()
show("forward (march..feat + chain)", e0.elapsed_time(e1))
eng.zero_grad(); dbg.zero_()
e0.record(); eng.loss_backward(tgt); e1.record(); torch.cuda.synchronize()
check(lib.fgs_debug_gemm_phases(None))
show("loss + backward (chain + dW + scatter)", e0.elapsed_time(e1))
show("  dW kernel (slot 8 = slabs of CTA 0)", e0.elapsed_time(e1), 10)
d = dbg.cpu().tolist()
print("  dW extra: producer wait raw ring %.1f us, wait stage free %.1f us, convert+store %.1f us; issuer inside issue_a %.1f us" %
      (d[12] / 1965, d[13] / 1965, d[14] / 1965, d[16] / 1965))
print("  bwd planes extra: issuer inside issue() %.1f us; consumer fence+colsum flush %.1f us" % (d[8] / 1965, d[24] / 1965))
