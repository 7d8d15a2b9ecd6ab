"""Generates tests/golden/*.npz by EXECUTING THE REFERENCE's own functions (imported from
/root/reference, never copied) on small seeded inputs.  Run in the build container only:

    python tests/golden/make_golden.py

The GPU box has no /root/reference; tests read the committed .npz files.
"""
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    sys.path.insert(0, REF)
    from xKV.customized_cache.fake_layer_merge_dynamic_cache import (  # reference code, executed as-is
        fake_minicache_merge,
        fake_svd,
        slerp_merge_rows_batch,
    )
    from transformers.models.mistral.modeling_mistral import apply_rotary_pos_emb

    torch.set_num_threads(1)
    out = {}
    # --- fake_svd cases: (bs, heads_total, seq, head_dim), rank ---
    cases = {
        "svd_a": ((1, 8, 48, 16), 8),      # 2 layers x 4 heads, rank 8
        "svd_b": ((2, 4, 40, 8), 5),       # batch 2 (batched SVD), odd rank
        "svd_c": ((1, 4, 12, 16), 64),     # rank >= min(m, n): slicing is a no-op
        "svd_d": ((1, 16, 96, 32), 64),    # 4 layers x 4 heads
    }
    for name, (shape, rank) in cases.items():
        g = torch.Generator().manual_seed(hash(name) % 1000 + 7)
        base = torch.randn(*shape, generator=g)
        # give the matrix a decaying spectrum so truncation is meaningful
        bs, nh, sl, hd = shape
        x = base.transpose(1, 2).reshape(bs, sl, nh * hd)
        u, s, vh = torch.linalg.svd(x, full_matrices=False)
        s = s * torch.arange(1, s.shape[-1] + 1, dtype=torch.float32) ** -0.7
        x = (u * s[:, None, :]) @ vh
        inp = x.view(bs, sl, nh, hd).transpose(1, 2).contiguous()
        ref = fake_svd(inp, rank)
        out[f"{name}_in"] = inp.numpy()
        out[f"{name}_rank"] = np.array(rank)
        out[f"{name}_out"] = ref.contiguous().numpy()
    # --- RoPE (transformers' function the reference imports at cache:5-7) ---
    g = torch.Generator().manual_seed(11)
    k = torch.randn(1, 4, 24, 16, generator=g)
    inv = 1.0 / (10000.0 ** (torch.arange(0, 16, 2).float() / 16))
    ang = torch.outer(torch.arange(24).float(), inv)
    emb = torch.cat([ang, ang], -1)
    cos, sin = emb.cos()[None], emb.sin()[None]
    _, k_rot = apply_rotary_pos_emb(k, k, cos, sin)
    out["rope_k"], out["rope_cos"], out["rope_sin"], out["rope_out"] = k.numpy(), cos.numpy(), sin.numpy(), k_rot.numpy()
    # --- SLERP / MiniCache branch ---
    g = torch.Generator().manual_seed(13)
    x1 = torch.randn(64, 16, generator=g)
    x2 = x1 + 0.3 * torch.randn(64, 16, generator=g)
    x2[5] = x1[5] * 2.0  # a parallel row -> linear fallback
    e, dm, n1, n2 = slerp_merge_rows_batch(x1, x2, t=0.6, gamma=0.05)
    e1, e2 = fake_minicache_merge(x1, x2, t=0.6, gamma=0.05)
    out.update(slerp_x1=x1.numpy(), slerp_x2=x2.numpy(), slerp_e=e.numpy(), slerp_mask=dm.numpy(),
               slerp_e1=e1.numpy(), slerp_e2=e2.numpy())
    np.savez_compressed(os.path.join(HERE, "reference_vectors.npz"), **out)
    print("wrote", os.path.join(HERE, "reference_vectors.npz"), {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
