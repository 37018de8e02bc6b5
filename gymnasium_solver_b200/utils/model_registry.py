"""model_id presets (reference: utils/model_registry.py:17-76).  ``mlp_64x64`` is new: BASELINE.json's C2 config names a
64x64 MLP that the reference registry does not define.  CNN presets are out of scope for this engine."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Dict, Tuple


@dataclass
class ModelSpec:
    policy: str
    hidden_dims: Tuple[int, ...]
    activation: str = "relu"
    policy_kwargs: Dict[str, Any] = field(default_factory=dict)


MODEL_REGISTRY: Dict[str, ModelSpec] = {
    "mlp_tiny": ModelSpec("mlp_actorcritic", (64,)),
    "mlp_64x64": ModelSpec("mlp_actorcritic", (64, 64)),
    "mlp_small": ModelSpec("mlp_actorcritic", (128, 128)),
    "mlp_medium": ModelSpec("mlp_actorcritic", (256, 256)),
    "mlp_large": ModelSpec("mlp_actorcritic", (512, 512)),
}


def resolve_model_spec(model_id: str) -> ModelSpec:
    assert model_id in MODEL_REGISTRY, f"Unknown model_id: '{model_id}'. Available models: {sorted(MODEL_REGISTRY)}"
    return MODEL_REGISTRY[model_id]


def list_models() -> list[str]:
    return sorted(MODEL_REGISTRY)
