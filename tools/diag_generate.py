import sys; sys.path.insert(0, "/root/repo")
import torch
from transformers import LlamaConfig, LlamaForCausalLM
from tests.oracle_cache import OracleCache
from xkv_b200.configurations import generate_consecutive_xKV_config
from xkv_b200.customized_cache import FakeLayerMergingCache
from xkv_b200.patch import KVCompress
from tests.test_generate_gpu import _decode_logits

mc = LlamaConfig(hidden_size=4096, intermediate_size=1024, num_hidden_layers=8, num_attention_heads=32, num_key_value_heads=8,
                 head_dim=128, vocab_size=1024, max_position_embeddings=16384, rope_theta=500000.0)
mc._attn_implementation = "sdpa"
torch.manual_seed(0)
model = LlamaForCausalLM(mc).to(device="cuda", dtype=torch.bfloat16).eval()
cfg = generate_consecutive_xKV_config(num_layers=8, end_layer=-1, group_size=4, rank_k=512, rank_v=768)
KVCompress(xKV_config=cfg)(model)
for name, ids in [("iid tokens", torch.randint(0, 1024, (1, 8192), device="cuda")),
                  ("48 distinct tokens", torch.randint(0, 48, (1, 8192), device="cuda"))]:
    for layer in model.model.layers:
        layer.self_attn.xkv_fused_decode = True
    c1 = FakeLayerMergingCache(cfg)
    lg1, _ = _decode_logits(model, c1, ids, steps=3)
    for layer in model.model.layers:
        layer.self_attn.xkv_fused_decode = False
    c2 = OracleCache(cfg)
    lg2, _ = _decode_logits(model, c2, ids, steps=3)
    # dense (uncompressed) baseline: plain DynamicCache semantics through the oracle with huge rank = no-op
    cfg0 = generate_consecutive_xKV_config(num_layers=8, end_layer=-1, group_size=4, rank_k=100000, rank_v=100000)
    c3 = OracleCache(cfg0)
    lg3, _ = _decode_logits(model, c3, ids, steps=3)
    k1, v1 = c1.materialize(0)
    k2, v2 = c2.layers[0].keys, c2.layers[0].values
    k3, v3 = c3.layers[0].keys, c3.layers[0].values
    def rel(a, b): return ((a.float() - b.float()).norm() / b.float().norm()).item()
    print(name, "| logits ours-oracle", (lg1[1:] - lg2[1:]).abs().max().item(), "ours-dense", (lg1[1:] - lg3[1:]).abs().max().item(),
          "oracle-dense", (lg2[1:] - lg3[1:]).abs().max().item(), "scale", lg3[1:].abs().max().item())
    print("   layer0 K: err ours", rel(k1[:, :, :8192], k3[:, :, :8192]), "oracle", rel(k2[:, :, :8192], k3[:, :, :8192]), "ours vs oracle", rel(k1[:, :, :8192], k2[:, :, :8192]))
    print("   layer0 V: err ours", rel(v1[:, :, :8192], v3[:, :, :8192]), "oracle", rel(v2[:, :, :8192], v3[:, :, :8192]), "ours vs oracle", rel(v1[:, :, :8192], v2[:, :, :8192]), flush=True)
