"""fgs_b200: B200-native (sm_100a) implementation of FGS-NeRF's per-ray training/rendering hot
path, behind the reference's own operator API.

Submodules (imported lazily -- importing any of them loads libfgs_b200.so and fails loudly if
it has not been built):
  render_utils_cuda, adam_upd_cuda, total_variation_cuda   module-level ops (reference names)
  grid (DenseGrid, MaskGrid, create_grid), adam (MaskedAdam), nerf (nerf, Alphas2Weights, MaskCache)
  engine (FineStepEngine: fused fwd+bwd+TV+Adam step), scene (synthetic scenes of SURVEY 8d)
"""
__version__ = "0.1.0"
