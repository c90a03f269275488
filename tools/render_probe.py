"""GPU probe: where a full-image render (BASELINE config 5) spends its time -- host enqueue time against device time,
and the per-launch device times of one 8192-ray chunk."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fgs_nerf_b200 import engine, scene, _lib
G = int(sys.argv[1]) if len(sys.argv) > 1 else 384
dev = torch.device("cuda:0")
sc = scene.make_scene(G=G)
eng = engine.FineStepEngine(sc, 8192, dev)
R = 800 * 800
ro, rd, vd = [t.to(dev) for t in scene.make_rays(N=R)[:3]]
for _ in range(2):
    eng.render(ro, rd, vd, global_step=5000)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
out = eng.render(ro, rd, vd, global_step=5000)
e1.record()
t_host = time.perf_counter() - t0          # includes render()'s final sync on the overflow word
torch.cuda.synchronize()
print(f"grid {G}^3: render of {R} rays: device {e0.elapsed_time(e1):.2f} ms, host call {1e3 * t_host:.2f} ms")
# host-only cost: enqueue a chunk forward repeatedly without waiting
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(79):
    eng.forward(ro[:8192], rd[:8192], vd[:8192], global_step=5000, render_depth=True, render_grad=True)
t_enq = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print(f"79 chunk forwards: host enqueue {1e3 * t_enq:.2f} ms, until done {1e3 * t_all:.2f} ms")
_lib.prof_begin()
eng.forward(ro[8192:16384].contiguous(), rd[8192:16384].contiguous(), vd[8192:16384].contiguous(), global_step=5000, render_depth=True, render_grad=True)
rows = _lib.prof_end()
print("one chunk:", eng.counters())
for n, ms in rows:
    print(f"   {n:28s} {1e3 * ms:8.1f} us")
print(f"   sum {1e3 * sum(ms for _, ms in rows):.1f} us")
