"""mvd_b200 — B200-native (sm_100a) implementation of MVD's multi-view denoising hot path."""
__version__ = "0.1.0"
