"""Rebuilds the inputs of tests/golden/fine_golden.npz from seeds (see make_fine_golden.py)."""
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fine_golden.npz")


def load_golden():
    return np.load(GOLD)


def golden_inputs(gold=None):
    from fgs_nerf_b200 import scene
    gold = gold if gold is not None else load_golden()
    G, N = int(gold["meta_G"]), int(gold["meta_N"])
    sc = scene.make_scene(G=G, mask_G=int(gold["meta_mask_G"]))
    rays = scene.make_rays(N=N, H=int(gold["meta_H"]), W=int(gold["meta_W"]), n_views=int(gold["meta_n_views"]))
    return sc, rays, int(gold["meta_global_step"])


def dense_k0_grad(gold, shape):
    g = np.zeros(int(np.prod(shape)), np.float32)
    g[gold["grad_k0_idx"]] = gold["grad_k0_val"]
    return g.reshape(shape)
