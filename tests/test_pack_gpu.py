"""GPU parity of the gather kernel against torch.cat + transpose + reshape (bit-exact).

Reference semantics: fake_layer_merge_dynamic_cache.py:170-171 (cat over heads) and :13-14
(token-major reshape)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref_pack(layers):
    cat = torch.cat(layers, dim=1)
    bs, nh, sl, hd = cat.shape
    return cat.transpose(1, 2).reshape(bs, sl, nh * hd)


@pytest.mark.parametrize(
    "g,bs,h,s,d,token_major",
    [
        (4, 1, 8, 257, 128, True),   # HF layout: (bs, S, H, D) viewed as (bs, H, S, D)
        (4, 1, 8, 257, 128, False),  # contiguous (bs, H, S, D)
        (1, 2, 8, 64, 128, False),   # single-layer group, batch 2
        (3, 1, 1, 100, 512, True),   # MLA latent slot: one head of 512
        (2, 1, 2, 1, 64, False),     # one token
        (7, 1, 8, 33, 128, True),    # 7-layer group (configs/grouped_layers.yaml)
    ],
)
def test_pack_matches_cat_reshape(g, bs, h, s, d, token_major):
    from xkv_b200 import ops

    torch.manual_seed(0)
    layers = []
    for _ in range(g):
        if token_major:
            t = torch.randn(bs, s, h, d, device="cuda").bfloat16().transpose(1, 2)
        else:
            t = torch.randn(bs, h, s, d, device="cuda").bfloat16()
        layers.append(t)
    x = ops.pack_group(layers)
    torch.cuda.synchronize()
    assert torch.equal(x, _ref_pack(layers))
    # round trip through the inverse scatter
    outs = [torch.zeros_like(t) for t in layers]
    ops.unpack_group(x, outs)
    torch.cuda.synchronize()
    for a, b in zip(outs, layers):
        assert torch.equal(a, b)


def test_pack_empty_prefill():
    from xkv_b200 import ops

    layers = [torch.empty(1, 8, 0, 128, device="cuda", dtype=torch.bfloat16) for _ in range(4)]
    x = ops.pack_group(layers)
    assert x.shape == (1, 0, 4096)


def test_pack_rejects_cpu_tensors():
    from xkv_b200 import _lib, ops

    with pytest.raises(_lib.XkvError):
        ops.pack_group([torch.zeros(1, 1, 8, 8, dtype=torch.bfloat16)])
