"""Same module name and functions as the reference's root ``utils.py`` (``add_common_args``,
``apply_kv_compress_patch``, ``load_model_and_tokenizer``): the reference's driver scripts import these by
``from utils import ...``.  The implementation lives in ``xkv_b200/cli.py``."""
from xkv_b200.cli import add_common_args, apply_kv_compress_patch, config_from_args, load_model_and_tokenizer  # noqa: F401
