"""Fused fine-stage step (placeholder until the fused kernels land in this round)."""
