#!/usr/bin/env python
"""Does the size of the ray pool change the iteration time?  bench.py rotates 4 resident batches; real training (and
FineTrainer) draws a new random batch every iteration, so the k0 voxels an iteration touches (p, m, v, g: 64 B per
channel-voxel) are not the ones the previous iterations left in the 126 MB L2.  Times FineTrainer.run on pools of
4, 16 and 64 batches of 8192 rays (fresh engine each, same scene), CUDA events, 30 iterations after 5 warm-ups."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from fgs_nerf_b200 import engine, scene, training
    dev = torch.device("cuda:0")
    N, G, K = 8192, 160, 30
    sc = scene.make_scene(G=G, mask_G=120)
    rays_all = [t.to(dev) for t in scene.make_rays(N=64 * N)]
    out = []
    for pool in (4, 16, 64):
        rays = [t[:pool * N].contiguous() for t in rays_all]
        eng = engine.FineStepEngine(sc, N, dev)
        eng.set_nonempty_mask()
        tr = training.FineTrainer(eng, *rays, cfg_train=dict(scene.FINE_TRAIN, N_rand=N, lrate_decay=20), start=5000,
                                  rng=np.random.RandomState(0))
        tr.run(5)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        tr.run(K)
        b.record()
        b.synchronize()
        ms = a.elapsed_time(b) / K
        cnt = eng.counters()
        out.append(dict(pool_batches=pool, ms_per_step=round(ms, 4), rays_per_s=round(N / ms * 1e3), M2_last=cnt["M2"]))
        print(json.dumps(out[-1]), flush=True)
        del eng, tr
    if len(sys.argv) > 1:
        json.dump(out, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
