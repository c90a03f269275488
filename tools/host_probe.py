"""How long does the HOST take to enqueue one training step (no synchronisation inside)?"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fgs_nerf_b200 import engine, scene
dev = torch.device("cuda:0")
sc = scene.make_scene(G=160)
ro, rd, vd, tgt = [t.to(dev) for t in scene.make_rays(N=8192)]
eng = engine.FineStepEngine(sc, 8192, dev)
eng.set_nonempty_mask()
for i in range(10):
    eng.train_step(ro, rd, vd, tgt, 5000 + i)
torch.cuda.synchronize()
K = 60
t0 = time.perf_counter()
for i in range(K):
    eng.train_step(ro, rd, vd, tgt, 5010 + i)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1e3*(t1-t0)/K:.3f} ms/step; total incl. drain {1e3*(t2-t0)/K:.3f} ms/step")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for i in range(20):
    eng.train_step(ro, rd, vd, tgt, 5100 + i)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
