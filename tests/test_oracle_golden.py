"""CPU: the oracle (oracle/xkv_oracle.py) against golden vectors produced by the reference's own
functions (tests/golden/make_golden.py imports /root/reference/xKV/... and executes them)."""
import os

import numpy as np
import pytest
import torch

from oracle import xkv_oracle as O

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.npz"))


@pytest.mark.parametrize("name", ["svd_a", "svd_b", "svd_c", "svd_d"])
def test_fake_svd_matches_reference_outputs(name):
    x = torch.from_numpy(GOLD[f"{name}_in"])
    rank = int(GOLD[f"{name}_rank"])
    ref = torch.from_numpy(GOLD[f"{name}_out"])
    got = O.fake_svd(x, rank)
    assert got.shape == ref.shape
    # same LAPACK driver, possibly different threading: equal to fp32 round-off
    assert torch.allclose(got, ref, rtol=1e-4, atol=1e-5)


def test_fake_svd_rank_beyond_min_dim_is_identity():
    x = torch.from_numpy(GOLD["svd_c_in"])
    got = O.fake_svd(x, int(GOLD["svd_c_rank"]))
    assert torch.allclose(got, x, atol=1e-5)


def test_rope_matches_transformers_function_the_reference_imports():
    k = torch.from_numpy(GOLD["rope_k"])
    got = O.apply_rope(k, torch.from_numpy(GOLD["rope_cos"]), torch.from_numpy(GOLD["rope_sin"]))
    assert torch.allclose(got, torch.from_numpy(GOLD["rope_out"]), atol=1e-6)


def test_slerp_branch_matches_reference_outputs():
    x1, x2 = torch.from_numpy(GOLD["slerp_x1"]), torch.from_numpy(GOLD["slerp_x2"])
    e, mask, _, _ = O.slerp_merge_rows_batch(x1, x2, t=0.6, gamma=0.05)
    assert torch.equal(mask, torch.from_numpy(GOLD["slerp_mask"]))
    assert torch.allclose(e, torch.from_numpy(GOLD["slerp_e"]), atol=1e-6, equal_nan=True)
    e1, e2 = O.fake_minicache_merge(x1, x2, t=0.6, gamma=0.05)
    assert torch.allclose(e1, torch.from_numpy(GOLD["slerp_e1"]), atol=1e-6, equal_nan=True)
    assert torch.allclose(e2, torch.from_numpy(GOLD["slerp_e2"]), atol=1e-6, equal_nan=True)


def test_merge_group_column_order_and_dtype_roundtrip():
    """cache:170-182: layers are concatenated over heads, so the factorised matrix's columns are ordered
    (layer, head, dim); outputs come back per layer in the cache dtype."""
    torch.manual_seed(0)
    keys = [torch.randn(1, 2, 20, 8).bfloat16() for _ in range(3)]
    vals = [torch.randn(1, 2, 20, 8).bfloat16() for _ in range(3)]
    k_hat, v_hat = O.merge_group(keys, vals, rank_k=4, rank_v=48)
    assert len(k_hat) == 3 and all(k.shape == (1, 2, 20, 8) and k.dtype == torch.bfloat16 for k in k_hat)
    # rank 48 >= min(20, 48): values survive up to the bf16 round trip of the fp32 SVD product
    for v, vh in zip(vals, v_hat):
        assert torch.allclose(v.float(), vh.float(), atol=2e-2)
    # the rank-4 key reconstruction is the best rank-4 approximation of the stacked matrix
    x = torch.cat(keys, dim=1).float().transpose(1, 2).reshape(20, 48)
    xh = torch.cat(k_hat, dim=1).float().transpose(1, 2).reshape(20, 48)
    s = torch.linalg.svdvals(x)
    best = (s[4:] ** 2).sum().sqrt()
    assert abs((x - xh).norm().item() - best.item()) < 0.05 * best.item()
    # merge_key=False leaves keys untouched (cache:175)
    k_same, _ = O.merge_group(keys, vals, rank_k=4, rank_v=4, merge_key=False)
    assert all(torch.equal(a, b) for a, b in zip(keys, k_same))


def test_decode_attention_appends_uncompressed_token():
    torch.manual_seed(1)
    q = torch.randn(1, 8, 1, 16)
    kc, vc = torch.randn(1, 2, 10, 16), torch.randn(1, 2, 10, 16)
    kn, vn = torch.randn(1, 2, 1, 16), torch.randn(1, 2, 1, 16)
    out = O.decode_attention(q, kc, vc, kn, vn)
    # manual GQA softmax
    k = torch.cat([kc, kn], 2).repeat_interleave(4, 1)
    v = torch.cat([vc, vn], 2).repeat_interleave(4, 1)
    p = torch.softmax(q @ k.transpose(-1, -2) / 4.0, -1)
    assert torch.allclose(out, p @ v, atol=1e-5)
