"""Effective hyper-parameters of the reference model for a given YAML dict.

The reference reads most keys from the TOP level of the config dict while the shipped YAMLs nest them under
`model:` (reference src/model.py:803,822,835-837,854-862,951 vs configs/experiment_B.yaml:9-11,25-28,75-77), so
every shipped config builds the same network.  This module reproduces those lookups exactly — it does not
"fix" them — and is unit-tested against attributes of the real reference model
(tests/golden/effective_config.json).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

DEFAULT_COGNITIVE_MODULES = ["ambient_stream", "iterative_focal_stream", "exif_prior_database"]  # demo.py:46-52


@dataclass(frozen=True)
class EffectiveConfig:
    backbone_size: str
    feature_dim: int
    use_lora: bool
    lora_rank: int
    use_ambient: bool
    use_focal: bool
    use_iterative: bool
    use_exif: bool
    num_cameras: int
    curiosity_guided: bool
    num_iterations: int
    focus_strength: float
    focal_hidden_dim: int
    enable_hierarchical_curiosity: bool
    lora_merge_target: Optional[str] = None  # NOT a reference key: see model.CognitiveAimModel._lora_delta
    lora_mode: str = "merge"  # NOT a reference key: "merge" (into the packed weight) or "fused" (extra K columns of the GEMM)
    fusion_dim: int = 192  # src/model.py:904-905


def effective_config(config: dict, camera_info: Optional[dict] = None) -> EffectiveConfig:
    """Mirror of the lookups in CognitiveAimModel.__init__ (reference src/model.py:798-952)."""
    backbone_size = config.get("backbone_size", "base")  # :803 (top level!)
    feature_dim = 1024 if backbone_size == "large" else 768  # :804-812
    model_cfg = config.get("model", {}) or {}
    modules = model_cfg.get("cognitive_modules", config.get("cognitive_modules", []))  # :835-836
    cga = config.get("curiosity_guided_attention", {}) or {}  # :854 (top level; key is `enabled`)
    focal_cfg = config.get("focal_config", {}) or {}  # :855 (top level)
    use_iter = "iterative_focal_stream" in modules
    use_focal = use_iter or ("focal_stream" in modules)
    return EffectiveConfig(
        backbone_size="large" if backbone_size == "large" else "base",
        feature_dim=feature_dim,
        use_lora=bool(config.get("use_lora", False)),  # :822
        lora_rank=int(config.get("lora_rank", 16)),
        use_ambient="ambient_stream" in modules,
        use_focal=use_focal,
        use_iterative=use_iter,
        use_exif=("exif_prior_database" in modules) and bool(camera_info),  # :881
        num_cameras=int(camera_info["num_cameras"]) if camera_info else 0,
        curiosity_guided=bool(cga.get("enabled", False)),
        num_iterations=int(focal_cfg.get("num_iterations", 3)),  # :860
        focus_strength=float(focal_cfg.get("focus_strength", 1.5)),  # :862
        focal_hidden_dim=int(config.get("focal_hidden_dim", 256)),  # :859
        enable_hierarchical_curiosity=bool(config.get("enable_hierarchical_curiosity", True)),  # :951
        lora_merge_target=config.get("lora_merge_target"),
        lora_mode=str(config.get("lora_mode", "merge")),
    )
