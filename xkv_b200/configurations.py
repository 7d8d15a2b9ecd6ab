"""Merge configuration of the xKV path: which layers form a group and at what rank.

Host-side mirror of the reference's ``xKV/configurations.py`` (same public names, fields, defaults,
YAML schema and error behaviour, so reference YAML files load unchanged):

* ``LayerGroup``                          reference configurations.py:27-50
* ``xKVConfig``                           :53-250  (``from_yaml`` :162-189, ``to_yaml`` :209-231,
                                          ``get_group_for_layer`` :154-160)
* ``generate_consecutive_layer_groups``   :254-273
* ``generate_consecutive_xKV_config``     :276-323

Nothing here touches the GPU; it is the configuration surface the north-star asks to keep.
"""
from __future__ import annotations

import logging
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import yaml

logger = logging.getLogger("xkv_b200")

_MERGE_IMPLS = ("svd", "slerp")
_YAML_ROOT = "xKV_config"


@dataclass
class LayerGroup:
    """A set of layers whose K (and V) are compressed jointly.

    Only the parameters of the active ``layer_merge_impl`` are meaningful once the owning
    :class:`xKVConfig` has been constructed: the others are reset to ``None``."""

    layers: List[int] = field(default_factory=list)
    rank_k: Optional[int] = None
    rank_v: Optional[int] = None
    slerp_t: Optional[float] = None
    slerp_gamma: Optional[float] = None

    def __post_init__(self):
        if len(self.layers) == 0:
            raise ValueError("LayerGroup must have at least one layer index.")


@dataclass
class xKVConfig:  # noqa: N801  (name kept from the reference API)
    """Global defaults plus the list of layer groups.

    Per-group values win over the global ones; after construction every group carries definite
    ``rank_k``/``rank_v`` (svd) or ``slerp_t``/``slerp_gamma`` (slerp)."""

    num_layers: Optional[int] = None
    layer_merge_impl: str = "svd"
    rank_k: Optional[int] = None
    rank_v: Optional[int] = None
    slerp_t: float = 0.5
    slerp_gamma: float = 1.0
    merge_key: bool = True
    merge_value: bool = True
    layer_groups: List[LayerGroup] = field(default_factory=list)
    extra_kwargs: dict = field(default_factory=dict)
    _layer_map: Dict[int, LayerGroup] = field(init=False, default_factory=dict)

    def __post_init__(self):
        if self.layer_merge_impl not in _MERGE_IMPLS:
            raise ValueError(
                f"Invalid layer_merge_impl '{self.layer_merge_impl}'. Must be 'svd' or 'slerp'."
            )
        use_svd = self.layer_merge_impl == "svd"
        for group in self.layer_groups:
            if use_svd:
                group.rank_k = self.rank_k if group.rank_k is None else group.rank_k
                group.rank_v = self.rank_v if group.rank_v is None else group.rank_v
                group.slerp_t = group.slerp_gamma = None
            else:
                group.slerp_t = self.slerp_t if group.slerp_t is None else group.slerp_t
                group.slerp_gamma = self.slerp_gamma if group.slerp_gamma is None else group.slerp_gamma
                group.rank_k = group.rank_v = None
        self._layer_map = self._build_layer_to_group_map(raise_if_duplicate=True)
        if self.num_layers is not None:
            self._validate_num_layers()

    # ---- validation / lookup -----------------------------------------------------------------
    def _validate_num_layers(self) -> None:
        for group in self.layer_groups:
            for layer in group.layers:
                if layer >= self.num_layers:
                    raise ValueError(
                        f"Group has a layer index {layer} which exceeds the declared "
                        f"num_layers={self.num_layers} (max index {self.num_layers - 1})."
                    )

    def _build_layer_to_group_map(self, raise_if_duplicate: bool = True) -> Dict[int, LayerGroup]:
        mapping: Dict[int, LayerGroup] = {}
        for group in self.layer_groups:
            for layer in group.layers:
                if raise_if_duplicate and layer in mapping:
                    raise ValueError(f"Layer {layer} appears in multiple groups: {mapping[layer]} and {group}")
                mapping[layer] = group
        return mapping

    def get_group_for_layer(self, layer_idx: int) -> Optional[LayerGroup]:
        """The group containing ``layer_idx``, or ``None`` for an un-grouped layer."""
        return self._layer_map.get(layer_idx)

    # ---- YAML round trip ---------------------------------------------------------------------
    @classmethod
    def from_yaml(cls, path: str) -> "xKVConfig":
        """Load ``{xKV_config: {<global fields>, layer_groups: [{layers: [...], rank_k: ...}, ...]}}``."""
        with open(path, "r") as handle:
            document = yaml.safe_load(handle) or {}
        body = dict(document.get(_YAML_ROOT, {}))
        groups = [LayerGroup(**entry) for entry in body.pop("layer_groups", [])]
        return cls(layer_groups=groups, **body)

    def to_dict(self) -> dict:
        """Top-level fields (``layer_groups`` excluded), with ``extra_kwargs`` flattened in."""
        fields = {
            name: getattr(self, name)
            for name in ("num_layers", "layer_merge_impl", "rank_k", "rank_v", "slerp_t", "slerp_gamma",
                         "merge_key", "merge_value")
        }
        fields.update(self.extra_kwargs)
        return fields

    def to_yaml(self, path: str) -> None:
        body = self.to_dict()
        body["layer_groups"] = []
        for group in self.layer_groups:
            entry = {"layers": group.layers}
            for name in ("rank_k", "rank_v", "slerp_t", "slerp_gamma"):
                value = getattr(group, name)
                if value is not None:
                    entry[name] = value
            body["layer_groups"].append(entry)
        with open(path, "w") as handle:
            yaml.safe_dump({_YAML_ROOT: body}, handle, sort_keys=False)

    def __str__(self) -> str:
        head = [
            f"{type(self).__name__}(",
            "  # Global params:",
            f"  num_layers={self.num_layers},",
            f"  layer_merge_impl={self.layer_merge_impl!r},",
            f"  rank_k={self.rank_k!r}, rank_v={self.rank_v!r},",
            f"  slerp_t={self.slerp_t!r}, slerp_gamma={self.slerp_gamma!r},",
            f"  merge_key={self.merge_key}, merge_value={self.merge_value},",
            f"  # {len(self.layer_groups)} groups:",
        ]
        body = [f"    [{i}] -> {group!r}" for i, group in enumerate(self.layer_groups)]
        return "\n".join(head + body + [")"])


def generate_consecutive_layer_groups(start_layer: int, end_layer: int, group_size: int) -> List[LayerGroup]:
    """Chunk ``[start_layer, end_layer]`` (inclusive) into runs of ``group_size``; the last may be short."""
    return [
        LayerGroup(layers=list(range(first, min(first + group_size, end_layer + 1))))
        for first in range(start_layer, end_layer + 1, group_size)
    ]


def generate_consecutive_xKV_config(  # noqa: N802
    layer_merge_impl: str = "svd",
    start_layer: int = 0,
    end_layer: int = 31,
    num_layers: Optional[int] = None,
    group_size: int = 2,
    rank_k: Optional[int] = 256,
    rank_v: Optional[int] = 768,
    slerp_t: float = 0.5,
    slerp_gamma: float = 1.0,
    merge_key: bool = True,
    merge_value: bool = True,
    extra_kwargs: dict = None,
) -> xKVConfig:
    """Uniform xKV-``group_size`` configuration over consecutive layers (``end_layer=-1``: last layer)."""
    if end_layer == -1:
        assert num_layers is not None, "Must provide num_layers if end_layer is -1."
        logger.info("End layer not specified, using num_layer=%s - 1.", num_layers)
        end_layer = num_layers - 1
    return xKVConfig(
        num_layers=num_layers,
        layer_merge_impl=layer_merge_impl,
        rank_k=rank_k,
        rank_v=rank_v,
        slerp_t=slerp_t,
        slerp_gamma=slerp_gamma,
        merge_key=merge_key,
        merge_value=merge_value,
        layer_groups=generate_consecutive_layer_groups(start_layer, end_layer, group_size),
        extra_kwargs=extra_kwargs or {},
    )
