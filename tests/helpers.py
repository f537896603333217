"""Shared helpers for the parity tests: seeded synthetic inputs and error metrics."""
import math

import numpy as np
import torch
import torch.nn.functional as F


def round_to_bf16_(module: torch.nn.Module):
    """Make an fp32 oracle module carry exactly the weights the bf16 product sees."""
    with torch.no_grad():
        for p in module.parameters():
            p.copy_(p.to(torch.bfloat16).float())
    return module


def camera_matrix(azimuth_deg: float, elevation_deg: float = 20.0, radius: float = 1.8) -> torch.Tensor:
    """Look-at [3,4] world->camera matrix in the convention of the reference's create_camera_matrix
    (src/utils.py:51-85: columns right / up / -forward, translation = camera position)."""
    az, el = math.radians(azimuth_deg), math.radians(elevation_deg)
    pos = np.array([radius * math.cos(el) * math.sin(az), radius * math.sin(el), radius * math.cos(el) * math.cos(az)])
    fwd = -pos / np.linalg.norm(pos)
    right = np.cross(fwd, np.array([0.0, 1.0, 0.0]))
    right /= np.linalg.norm(right)
    up = np.cross(right, fwd)
    m = np.zeros((3, 4))
    m[:, 0], m[:, 1], m[:, 2], m[:, 3] = right, up, -fwd, pos
    return torch.from_numpy(m).float()


def synthetic_inputs(V: int, L: int, cfg: int, text_dim: int = 1024, text_len: int = 77):
    """Seeded inputs of SURVEY.md 8(d): latents, text, source latents, cameras, posenc projection."""
    g = lambda s: torch.Generator().manual_seed(s)  # noqa: E731
    latents = torch.randn(V, 4, L, L, generator=g(2))
    text = torch.randn(V * cfg, text_len, text_dim, generator=g(3))
    src_lat = 0.18215 * torch.randn(V, 4, L, L, generator=g(4))
    src_cam = torch.stack([camera_matrix(0.0)] * V)
    tgt_cam = torch.stack([camera_matrix(360.0 * i / V) for i in range(V)])
    proj = torch.randn(1024, 1020, generator=g(5)) / math.sqrt(1020)
    return dict(latents=latents, text=text, source_latents=src_lat, source_camera=src_cam, target_camera=tgt_cam,
                pos_proj=proj)


def metrics(got: torch.Tensor, ref: torch.Tensor):
    got, ref = got.detach().float().cpu().flatten(), ref.detach().float().cpu().flatten()
    max_abs = (got - ref).abs().max().item()
    ref_max = ref.abs().max().item()
    cos = F.cosine_similarity(got, ref, dim=0).item()
    return dict(max_abs=max_abs, ref_max=ref_max, rel=max_abs / max(1.0, ref_max), cos=cos)


def assert_close_bf16(got, ref, name, max_rel=2e-2, min_cos=0.999):
    """north_star tolerance for bf16 vs the fp32 oracle: cosine >= 0.999 and max-abs error <= 2e-2, the latter
    taken relative to max(1, max|ref|): one bf16 rounding of a value in [4, 8) is already 1.6e-2 absolute, so
    an absolute 2e-2 bound is only meaningful for O(1) activations."""
    m = metrics(got, ref)
    print(f"{name}: max_abs={m['max_abs']:.3e} ref_max={m['ref_max']:.2f} rel={m['rel']:.3e} cos={m['cos']:.6f}")
    assert m["cos"] >= min_cos, f"{name}: cosine {m['cos']} < {min_cos}"
    assert m["rel"] <= max_rel, f"{name}: normalised max-abs {m['rel']} > {max_rel}"
    return m
