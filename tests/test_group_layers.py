"""CPU: layer grouping from a CKA similarity matrix (reference group_layers.py:9-84) and the YAML it writes."""
import numpy as np
import torch

from xkv_b200 import group_layers as gl
from xkv_b200.configurations import xKVConfig


def _block_similarity(sizes, within=0.9, across=0.2, seed=0):
    n = sum(sizes)
    rng = np.random.default_rng(seed)
    sim = np.full((n, n), across) + 0.02 * rng.standard_normal((n, n))
    lo = 0
    for s in sizes:
        sim[lo:lo + s, lo:lo + s] = within + 0.02 * rng.standard_normal((s, s))
        lo += s
    return torch.tensor(0.5 * (sim + sim.T), dtype=torch.float32)


def test_block_similarity_recovers_contiguous_groups(tmp_path):
    sizes = [4, 4, 7, 5, 4, 8]
    sim = _block_similarity(sizes)
    groups = gl.group_layers(sim, ngroups=len(sizes))
    assert [len(g) for g in groups] == sizes
    assert sum(groups, []) == list(range(sum(sizes)))          # contiguous runs covering every layer once
    groups_w = gl.group_layers(sim, ngroups=len(sizes), weighted_by_index=True, index_scale=50.0)
    assert groups_w == groups
    path = tmp_path / "grouped_layers.yaml"
    torch.save(sim, tmp_path / "cka.pt")
    cfg = gl.main(["--cka_similarity", str(tmp_path / "cka.pt"), "--ngroups", str(len(sizes)), "--output_config", str(path)])
    loaded = xKVConfig.from_yaml(str(path))
    assert loaded.num_layers == 32 and [g.layers for g in loaded.layer_groups] == groups
    assert loaded.get_group_for_layer(9).layers == groups[2] and loaded.rank_k == cfg.rank_k == 512


def test_non_adjacent_members_of_a_cluster_become_separate_groups():
    # layers {0,1} and {4,5} resemble each other, {2,3} do not: one cluster, but two groups (the cache needs contiguity)
    sim = torch.full((6, 6), 0.1)
    for a in (0, 1, 4, 5):
        for b in (0, 1, 4, 5):
            sim[a, b] = 0.9
    sim[2, 3] = sim[3, 2] = 0.9
    groups = gl.group_layers(sim, ngroups=2)
    assert groups == [[0, 1], [2, 3], [4, 5]]


def test_layer_cka_is_one_for_rotated_copies_and_small_for_independent_layers():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(512, 64, generator=g)
    q = torch.linalg.qr(torch.randn(64, 64, generator=g))[0]
    layers = [x, 2.0 * x @ q, torch.randn(512, 64, generator=g)]
    cka = gl.layer_cka(layers)
    assert abs(cka[0, 1].item() - 1.0) < 1e-4 and cka[0, 2].item() < 0.3 and torch.allclose(cka, cka.t())
