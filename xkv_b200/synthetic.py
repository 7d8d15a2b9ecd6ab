"""Seeded synthetic KV for parity tests and bench.py.

A layer group's token-major matrix is drawn as  X = T diag(sigma) W^T  with a power-law spectrum
sigma_i = i^(-alpha) shared across the group's layers (that sharing is what cross-layer SVD exploits),
or i.i.d. N(0,1) (alpha=None, the worst case for any low-rank method), then cut back into per-layer
(1, H, S, D) bf16 tensors laid out like HuggingFace's K/V before caching (token-major memory viewed as
(bs, H, S, D)).  No dataset or checkpoint is involved.
"""
from __future__ import annotations

from typing import List, Optional

import torch


def group_matrix(tokens: int, cols: int, alpha: Optional[float], seed: int, device="cpu") -> torch.Tensor:
    """(tokens, cols) bf16 matrix with singular values ~ i^-alpha (alpha=None: i.i.d. Gaussian)."""
    g = torch.Generator(device=device).manual_seed(seed)
    if alpha is None:
        return torch.randn(tokens, cols, generator=g, device=device).to(torch.bfloat16)
    k = min(tokens, cols)
    t = torch.randn(tokens, k, generator=g, device=device)
    if tokens <= 8192:
        t = torch.linalg.qr(t)[0]
    else:  # tall Gaussian is orthonormal to within sqrt(k/tokens); good enough for a power-law spectrum
        t = t / tokens ** 0.5
    w = torch.linalg.qr(torch.randn(cols, k, generator=g, device=device))[0]
    s = torch.arange(1, k + 1, device=device, dtype=torch.float32) ** (-alpha)
    x = (t * s) @ w.t()
    x = x * (4.0 / x.abs().max())
    return x.to(torch.bfloat16)


def split_group(x: torch.Tensor, num_layers: int, heads: int, head_dim: int) -> List[torch.Tensor]:
    """Cut X (S, G*H*D) into G tensors (1, H, S, D) that are views of token-major (1, S, H, D) memory."""
    s = x.shape[0]
    xs = x.view(s, num_layers, heads, head_dim)
    return [xs[:, l].contiguous().view(1, s, heads, head_dim).transpose(1, 2) for l in range(num_layers)]


def make_group_kv(num_layers: int, heads: int, tokens: int, head_dim: int, alpha: Optional[float], seed: int,
                  device="cpu") -> List[torch.Tensor]:
    x = group_matrix(tokens, num_layers * heads * head_dim, alpha, seed, device)
    return split_group(x, num_layers, heads, head_dim)


def llama3_rope(tokens: int, head_dim: int = 128, theta: float = 500000.0, device="cpu", dtype=torch.bfloat16):
    """cos/sin tables (1, S, D) in the HF half-split convention for Llama-3.1 (llama3 scaling omitted:
    it only rescales frequencies and is irrelevant to kernel parity)."""
    inv = 1.0 / (theta ** (torch.arange(0, head_dim, 2, device=device, dtype=torch.float32) / head_dim))
    pos = torch.arange(tokens, device=device, dtype=torch.float32)
    ang = torch.outer(pos, inv)
    emb = torch.cat([ang, ang], dim=-1)
    return emb.cos()[None].to(dtype), emb.sin()[None].to(dtype)
