"""
Schema of the reference's transform configs (mirror of transform/transform_args.py, transform_scheme.py, transform_config.py).
Only the schema: `ModelCompressor.update_config` carries a model's transform_config through to config.json
(model_compressor.py:230-241).  Building and applying the transforms themselves (Hadamard / random rotations, fused weights) is a
separate subsystem of the reference and out of scope here.
"""
from __future__ import annotations

from enum import Enum

import torch
from pydantic import BaseModel, ConfigDict, Field, field_validator

from ..quantization.quant_args import TorchDtype

__all__ = ["TransformLocation", "TransformArgs", "TransformScheme", "TransformConfig"]


class TransformLocation(str, Enum):
    INPUT = "input"
    WEIGHT_INPUT = "weight_input"
    WEIGHT_OUTPUT = "weight_output"
    OUTPUT = "output"
    K_CACHE = "k_cache"
    Q_ATTN = "q_attn"

    def is_online(self) -> bool:
        """applied to activations at run time (everything except the two weight locations)"""
        return self not in (TransformLocation.WEIGHT_INPUT, TransformLocation.WEIGHT_OUTPUT)


class TransformArgs(BaseModel, use_enum_values=True):
    targets: list[str]
    location: TransformLocation
    inverse: bool = Field(default=False)
    ignore: list[str] = Field(default_factory=list)

    @field_validator("targets", "ignore", mode="before")
    @classmethod
    def _wrap(cls, v):
        return [v] if isinstance(v, str) else v

    def is_online(self) -> bool:
        return TransformLocation(self.location).is_online()

    model_config = ConfigDict(extra="forbid")


class TransformScheme(BaseModel):
    type: str
    apply: list[TransformArgs] = Field(default_factory=list)
    randomize: bool = Field(default=False)
    requires_grad: bool = Field(default=False)
    head_dim: int | None = Field(default=None)
    precision: TorchDtype = Field(default=torch.float32)

    model_config = ConfigDict(extra="forbid")


class TransformConfig(BaseModel):
    config_groups: dict[str, TransformScheme]

    model_config = ConfigDict(extra="forbid")

    def merge(self, other: "TransformConfig") -> None:
        """append the other config's groups under keys that do not collide"""
        for key, scheme in other.config_groups.items():
            name, i = key, 0
            while name in self.config_groups:
                i += 1
                name = f"{key}_{i}"
            self.config_groups[name] = scheme
