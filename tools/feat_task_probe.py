"""GPU probe: time of each independent row part of feat_kernel / feat_bwd_kernel on the bench workload."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fgs_nerf_b200 import engine, scene, _lib
from fgs_nerf_b200._lib import lib, check
dev = torch.device("cuda:0")
sc = scene.make_scene(G=160)
ro, rd, vd, tgt = [t.to(dev) for t in scene.make_rays(N=8192)]
eng = engine.FineStepEngine(sc, 8192, dev)
eng.set_nonempty_mask()
for i in range(3):
    eng.train_step(ro, rd, vd, tgt, 5000 + i)
torch.cuda.synchronize()
names = ["k0 | bwd k0", "encodings | bwd -", "reflection | bwd reflection", "taps d=0.5 | bwd z {0.5,1}", "taps d=1.0 | bwd z {1.5,2}", "taps d=1.5 | bwd y {0.5,1}", "taps d=2.0 | bwd y {1.5,2}", "- | bwd x {0.5,1}", "- | bwd x {1.5,2}", "all"]
for task in list(range(9)) + [-1]:
    lib.fgs_debug_feat_task(task)
    acc = {}
    for rep in range(5):
        _lib.prof_begin()
        eng.forward(ro, rd, vd, 5003, _skip_composite=True)
        eng.zero_grad()
        eng.loss_backward(tgt)
        for n, ms in _lib.prof_end():
            if n in ("fine feat", "fine feat_bwd"):
                acc.setdefault(n, []).append(ms)
        eng._grads_clean = False
        eng.zero_grad()
    print(f"{names[task]:26s} feat {1e3*min(acc['fine feat']):6.1f} us   feat_bwd {1e3*min(acc['fine feat_bwd']):6.1f} us   (M2 = {eng.counters()['M2']})")
lib.fgs_debug_feat_task(-1)
