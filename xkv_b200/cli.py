"""Command-line glue of the xKV patch (reference: ``utils.py:50-137`` at the repository root).

The flag names are part of the API surface ``BASELINE.json`` lists (``rank_k`` / ``rank_v`` /
``layer_group_size`` / ``start_layer_idx`` / ``end_layer_idx`` / ``customized_merge_config``), so they are kept
letter for letter, defaults included (``utils.py:96-137``):

    --model_name_or_path  --flash2  --xKV
    --rank_k 256  --rank_v 768  --layer_group_size 1
    --layer_merge_impl svd  --slerp_t 0.5  --slerp_gamma 0.05
    --merge_key  --merge_value          (store_true, default OFF; the README's ``--merge_k/--merge_v`` only work
                                         through argparse's prefix matching, which is therefore left enabled)
    --start_layer_idx 0  --end_layer_idx -1   (-1 = last layer, ``utils.py:86``)
    --customized_merge_config PATH      (a YAML merge configuration; wins over the flags, ``utils.py:70-72``)

``apply_kv_compress_patch(model, args)`` builds the configuration (YAML or consecutive groups over
``model.config.num_hidden_layers``) and installs the patch; like the reference it returns the patched model
(the reference's annotation promises a tuple, its body returns the model: ``utils.py:68,93``).
"""
from __future__ import annotations

import argparse
import logging
from typing import Any, Optional, Tuple

from .configurations import generate_consecutive_xKV_config, xKVConfig
from .patch import KVCompress

log = logging.getLogger("xkv_b200")

# (flag, kwargs) in the reference's order
_FLAGS = (
    ("--model_name_or_path", dict(type=str, help="model to load")),
    ("--flash2", dict(action="store_true", help="whether to use flash-attention2")),
    ("--xKV", dict(action="store_true", help="whether to enable xKV patch")),
    ("--rank_k", dict(type=int, default=256, help="Rank for SVD compression of keys")),
    ("--rank_v", dict(type=int, default=768, help="Rank for SVD compression of values")),
    ("--layer_group_size", dict(type=int, default=1,
                                help="The number of layers that will be grouped and decompose jointly")),
    ("--layer_merge_impl", dict(type=str, default="svd", help="The implementation for layer merge")),
    ("--slerp_t", dict(type=float, default=0.5, help="The interpolation ratio for SLERP")),
    ("--slerp_gamma", dict(type=float, default=0.05, help="The gamma for identifying divergent token in SLERP")),
    ("--merge_key", dict(action="store_true", help="Enable merging for keys")),
    ("--merge_value", dict(action="store_true", help="Enable merging for values")),
    ("--start_layer_idx", dict(type=int, default=0, help="The starting layer index for layer merging")),
    ("--end_layer_idx", dict(type=int, default=-1,
                             help="The ending layer index for layer merging. If -1, it will be the last layer.")),
    ("--customized_merge_config", dict(type=str, help="custom config file")),
)


def add_common_args(parser: argparse.ArgumentParser) -> argparse.ArgumentParser:
    """Register the reference's model / xKV flags on ``parser`` and return it (``utils.py:96-137``)."""
    for flag, kw in _FLAGS:
        parser.add_argument(flag, **kw)
    return parser


def config_from_args(args: argparse.Namespace, num_hidden_layers: int) -> xKVConfig:
    """The configuration ``apply_kv_compress_patch`` would install for ``args`` on a model of this depth."""
    if getattr(args, "customized_merge_config", None):
        log.info("Loading the customized merge config from %s", args.customized_merge_config)
        return xKVConfig.from_yaml(args.customized_merge_config)
    last = args.end_layer_idx if args.end_layer_idx != -1 else num_hidden_layers - 1
    return generate_consecutive_xKV_config(
        num_layers=num_hidden_layers, rank_k=args.rank_k, rank_v=args.rank_v, group_size=args.layer_group_size,
        layer_merge_impl=args.layer_merge_impl, slerp_t=args.slerp_t, slerp_gamma=args.slerp_gamma,
        merge_key=args.merge_key, merge_value=args.merge_value, start_layer=args.start_layer_idx, end_layer=last)


def apply_kv_compress_patch(model, args: argparse.Namespace, verbose: bool = True):
    """Install the xKV patch described by ``args`` on ``model`` and return the model (``utils.py:68-93``)."""
    config = config_from_args(args, model.config.num_hidden_layers)
    patch = KVCompress(xKV_config=config)
    if verbose:
        log.info("compression config: %s", patch.config)
    return patch(model)


def load_model_and_tokenizer(model_name_or_path: str, use_flash_attn2: bool = False) -> Tuple[Any, Any]:
    """bf16 model on the GPU with the sdpa (or FA2) attention implementation, in eval mode (``utils.py:50-66``)."""
    import torch
    from transformers import AutoModelForCausalLM, AutoTokenizer

    tokenizer = AutoTokenizer.from_pretrained(model_name_or_path, trust_remote_code=True)
    model = AutoModelForCausalLM.from_pretrained(
        model_name_or_path, dtype=torch.bfloat16, trust_remote_code=True, device_map="cuda",
        attn_implementation="flash_attention_2" if use_flash_attn2 else "sdpa")
    model.eval()
    return model, tokenizer


def main(argv: Optional[list] = None) -> int:
    """``python -m xkv_b200.cli --print-config ...``: show the merge configuration the flags describe."""
    parser = add_common_args(argparse.ArgumentParser(description=__doc__.split("\n")[0]))
    parser.add_argument("--num_hidden_layers", type=int, default=32, help="model depth used with --print-config")
    parser.add_argument("--print-config", action="store_true", help="print the resulting xKVConfig as YAML and exit")
    args = parser.parse_args(argv)
    cfg = config_from_args(args, args.num_hidden_layers)
    print(cfg)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
