"""Layer-group configuration from cross-layer similarity.

Counterpart of the reference's ``group_layers.py`` (repo root, :9-84): a (layers x layers) CKA similarity
matrix is clustered (average linkage on 1 - similarity, optionally damped by layer distance, :13-33), runs of
ADJACENT layers with equal label become groups (:50-56 — the cache assumes contiguous groups,
fake_layer_merge_dynamic_cache.py:161-165), and the result is written as an ``xKV_config`` YAML (:60-80) that
``xKVConfig.from_yaml`` loads.  Unlike the reference it can also measure the similarity itself from a KV cache
(``layer_cka``: linear CKA of the token-major layer matrices, on the device that holds them), so that the
notebook the reference relies on for that step (scripts/cka_similarity_analysis.ipynb) is not needed.

    python -m xkv_b200.group_layers --cka_similarity cka.pt --ngroups 8 --weighted_by_index \
        --output_config configs/grouped_layers.yaml
"""
from __future__ import annotations

import argparse
from typing import List, Optional, Sequence

import numpy as np
import torch

from .configurations import LayerGroup, xKVConfig


def layer_cka(layers: Sequence[torch.Tensor]) -> torch.Tensor:
    """Linear CKA between per-layer KV tensors (1, H, S, D) or matrices (S, n): CKA(X, Y) =
    ||Xc^T Yc||_F^2 / (||Xc^T Xc||_F ||Yc^T Yc||_F) with column-centred Xc, Yc (the mean-centring the reference's
    notebook applies).  Returns a (L, L) fp32 matrix on the layers' device."""
    mats = []
    for t in layers:
        x = t.transpose(1, 2).reshape(t.shape[2], -1) if t.dim() == 4 else t
        x = x.float()
        mats.append(x - x.mean(dim=0, keepdim=True))
    n = len(mats)
    self_norm = [torch.linalg.norm(m.t() @ m) for m in mats]
    out = torch.eye(n, dtype=torch.float32, device=mats[0].device)
    for i in range(n):
        for j in range(i + 1, n):
            v = torch.linalg.norm(mats[i].t() @ mats[j]) ** 2 / (self_norm[i] * self_norm[j])
            out[i, j] = out[j, i] = v
    return out


def _average_linkage(dissimilarity: np.ndarray, ngroups: int) -> np.ndarray:
    """Agglomerative clustering, average linkage, precomputed distances (what the reference asks sklearn for,
    group_layers.py:27-33); plain numpy so the tool has no dependency beyond it."""
    n = dissimilarity.shape[0]
    clusters = {i: [i] for i in range(n)}
    while len(clusters) > max(1, ngroups):
        keys = sorted(clusters)
        best, pair = None, None
        for ai, a in enumerate(keys):
            for b in keys[ai + 1:]:
                d = dissimilarity[np.ix_(clusters[a], clusters[b])].mean()
                if best is None or d < best:
                    best, pair = d, (a, b)
        a, b = pair
        clusters[a] = clusters[a] + clusters.pop(b)
    labels = np.empty(n, dtype=np.int64)
    for lab, members in enumerate(clusters[k] for k in sorted(clusters)):
        labels[members] = lab
    return labels


def group_layers(similarity, ngroups: int = 8, weighted_by_index: bool = False, index_scale: float = 50.0) -> List[List[int]]:
    """Cluster the layers and cut the label sequence into runs of adjacent equal labels."""
    sim = np.array(similarity.detach().cpu().numpy() if isinstance(similarity, torch.Tensor) else similarity, dtype=np.float64)
    if sim.ndim != 2 or sim.shape[0] != sim.shape[1]:
        raise ValueError("similarity must be a square (layers x layers) matrix")
    n = sim.shape[0]
    np.fill_diagonal(sim, 1.0)
    if weighted_by_index:
        dist = np.abs(np.subtract.outer(np.arange(n), np.arange(n)))
        sim = sim * np.exp(-dist / index_scale)
        sim = 0.5 * (sim + sim.T)
    labels = _average_linkage(1.0 - sim, ngroups)
    groups: List[List[int]] = []
    for layer in range(n):
        if layer == 0 or labels[layer] != labels[layer - 1]:
            groups.append([])
        groups[-1].append(layer)
    return groups


def config_from_groups(groups: Sequence[Sequence[int]], num_layers: int, rank_k: int = 512, rank_v: int = 768,
                       layer_merge_impl: str = "svd", slerp_t: float = 0.5, slerp_gamma: float = 0.05,
                       merge_key: bool = True, merge_value: bool = True) -> xKVConfig:
    return xKVConfig(num_layers=num_layers, layer_merge_impl=layer_merge_impl, rank_k=rank_k, rank_v=rank_v,
                     slerp_t=slerp_t, slerp_gamma=slerp_gamma, merge_key=merge_key, merge_value=merge_value,
                     layer_groups=[LayerGroup(layers=sorted(g), rank_k=rank_k, rank_v=rank_v) for g in groups])


def main(argv: Optional[Sequence[str]] = None) -> xKVConfig:
    ap = argparse.ArgumentParser(description="Cluster layers by CKA similarity into an xKV layer-group YAML config.")
    ap.add_argument("--cka_similarity", required=True, help=".pt file holding the (layers, layers) similarity matrix")
    ap.add_argument("--ngroups", type=int, default=8)
    ap.add_argument("--output_config", default="configs/grouped_layers.yaml")
    ap.add_argument("--weighted_by_index", action="store_true", help="damp the similarity by exp(-|i-j| / index_scale)")
    ap.add_argument("--index_scale", type=float, default=50.0)
    ap.add_argument("--rank_k", type=int, default=512)
    ap.add_argument("--rank_v", type=int, default=768)
    ap.add_argument("--verbose", action="store_true")
    args = ap.parse_args(argv)
    similarity = torch.load(args.cka_similarity, map_location="cpu")
    groups = group_layers(similarity, args.ngroups, args.weighted_by_index, args.index_scale)
    cfg = config_from_groups(groups, similarity.shape[0], args.rank_k, args.rank_v)
    if args.verbose:
        print("groups:", groups)
    cfg.to_yaml(args.output_config)
    print("Saving to:", args.output_config)
    return cfg


if __name__ == "__main__":
    main()
