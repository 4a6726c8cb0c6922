"""
ModelCompressor -- orchestration of compress / decompress over a whole model, same public surface
as compressors/model_compressors/model_compressor.py:31-273 of the reference
(from_pretrained_model, from_compression_config, compress_model, decompress_model, update_config,
the decompress-on-first-forward hook).

B200-native differences, invisible to callers:
  * single process: instead of one kernel sequence per module, all eligible modules of a device are
    compressed by ONE multi-tensor persistent launch (ops.batched -> ct_batched); the per-module
    plugin path is used for everything else (other formats, CPU-resident or odd-shaped tensors).
  * distributed (torch.distributed initialised, one rank per GPU): modules are dealt to ranks by the
    reference's greedy size-descending bin packing, each rank compresses only its own modules, and
    results are shared with dist.broadcast of the packed TENSORS (NCCL over NVLink) rather than
    pickled object lists.  No collective runs inside the compression itself.
"""
from __future__ import annotations

import json
import os
from functools import partial
from typing import Optional

import torch

from ... import __version__
from ...config import CompressionFormat
from ...quantization import QuantizationConfig, QuantizationStatus
from ...quantization.utils.helpers import is_module_quantized
from ..base import compress_module, decompress_module
from ..format import infer_model_format

__all__ = ["ModelCompressor"]

# keys of the config.json block (base.py of the reference)
QUANTIZATION_CONFIG_NAME = "quantization_config"
COMPRESSION_VERSION_NAME = "version"
QUANTIZATION_METHOD_NAME = "quant_method"
QUANTIZATION_METHOD = "compressed-tensors"
SPARSITY_CONFIG_NAME = "sparsity_config"
TRANSFORM_CONFIG_NAME = "transform_config"
CONFIG_NAME = "config.json"


class ModelCompressor:
    quantization_config: QuantizationConfig | None
    transform_config: object | None
    force_compression_format: CompressionFormat | None

    def __init__(self, quantization_config: Optional[QuantizationConfig] = None, transform_config=None,
                 force_compression_format: Optional[str] = None):
        self.quantization_config = quantization_config
        self.transform_config = transform_config
        self.force_compression_format = CompressionFormat(force_compression_format) if force_compression_format is not None else None

    # ---- constructors ---------------------------------------------------------------------
    @classmethod
    def from_compression_config(cls, compression_config):
        """entry used by the HF quantizer: an object carrying `.quantization_config` (a
        QuantizationConfig or its dict form) and optionally `.transform_config`"""
        q = getattr(compression_config, "quantization_config", None)
        if q is None and isinstance(compression_config, dict):
            q = compression_config.get(QUANTIZATION_CONFIG_NAME, compression_config)
        if q is None:
            raise ValueError(
                f"Support for compression config of type {type(compression_config)} is no longer supported. "
                "If you are attempting to use a Sparse24 model, note that the Sparse24 format is not longer "
                "supported by as of `compressed-tensors>0.14.0`"
            )
        if isinstance(q, dict):
            q = QuantizationConfig.model_validate(q)
        return cls(quantization_config=q, transform_config=getattr(compression_config, "transform_config", None))

    @classmethod
    def from_pretrained_model(cls, model: torch.nn.Module, sparsity_config_or_format=None, quantization_format: Optional[str] = None):
        quantization_config = QuantizationConfig.from_pretrained(model)
        if quantization_config is not None:
            quantization_config.format = infer_model_format(model, quantization_format).value
        return cls(quantization_config=quantization_config, transform_config=getattr(model, TRANSFORM_CONFIG_NAME, None),
                   force_compression_format=quantization_format)

    # ---- compression ------------------------------------------------------------------------
    def compress_model(self, model: torch.nn.Module, skip_compressed: bool = False, distributed: Optional[bool] = None,
                       recouple: bool = True, stats: Optional[dict] = None) -> None:
        """
        `distributed` (extension; default None = the reference's behaviour, follow `is_distributed()`): pass False
        when every rank holds its OWN model (independent replicas / shards), so that no module is dealt to another rank.
        `recouple` / `stats` (extensions, distributed path only): see `replace_module_parallel`.
        """
        modules = [
            m for _, m in model.named_modules(remove_duplicate=True)
            if is_module_quantized(m) and (not skip_compressed or getattr(m, "quantization_status", None) != QuantizationStatus.COMPRESSED)
        ]
        from ...distributed import is_distributed, replace_module_parallel

        from .batched import compress_modules_batched

        if distributed is None:
            distributed = is_distributed()
        if not distributed:
            compress_modules_batched(modules, self.force_compression_format)
        else:
            replace_module_parallel(modules, partial(compress_module, format=self.force_compression_format), desc=None,
                                    apply_many_fn=partial(compress_modules_batched, force_format=self.force_compression_format),
                                    recouple=recouple, stats=stats)
        if self.quantization_config is not None:
            self.quantization_config.quantization_status = QuantizationStatus.COMPRESSED
        self.add_decompress_hook(model)

    def decompress_model(self, model: torch.nn.Module, distributed: bool = False, recouple: bool = True, stats: Optional[dict] = None) -> None:
        """
        Default (`distributed=False`) = the reference: every rank decompresses every module it holds, locally, no collective
        (model_compressor.py:183-207; the decompress-on-first-forward hook always takes this path, so a rank-0-only forward
        cannot deadlock).  `distributed=True` is an explicit opt-in to the flow the reference leaves as a TODO (:196): modules
        are dealt to ranks exactly like in compress_model, each owner decompresses its share and the dense weights come back by
        NCCL broadcast.  It is a COLLECTIVE call: every rank must make it, holding the same compressed modules.
        """
        modules = [m for _, m in model.named_modules(remove_duplicate=True) if is_module_quantized(m)]
        from ...distributed import replace_module_parallel
        from .batched import decompress_modules_batched

        if not distributed:
            decompress_modules_batched(modules, self.force_compression_format)
        else:
            replace_module_parallel(modules, partial(decompress_module, format=self.force_compression_format), desc=None,
                                    apply_many_fn=partial(decompress_modules_batched, force_format=self.force_compression_format),
                                    recouple=recouple, stats=stats)
        if self.quantization_config is not None:
            self.quantization_config.quantization_status = QuantizationStatus.DECOMPRESSED
        self.remove_decompression_hook(model)

    # ---- config.json --------------------------------------------------------------------------
    def update_config(self, save_directory: str) -> None:
        if not any((self.quantization_config, self.transform_config)):
            return
        path = os.path.join(save_directory, CONFIG_NAME)
        data = {}
        if os.path.exists(path):
            with open(path, "r") as f:
                data = json.load(f)
        q = self.quantization_config.model_dump(exclude=["quant_method"], mode="json") if self.quantization_config is not None else {}
        t = self.transform_config.model_dump() if self.transform_config is not None else {}
        data[QUANTIZATION_CONFIG_NAME] = {
            COMPRESSION_VERSION_NAME: __version__,
            QUANTIZATION_METHOD_NAME: QUANTIZATION_METHOD,
            SPARSITY_CONFIG_NAME: {},
            TRANSFORM_CONFIG_NAME: t,
            **q,
        }
        with open(path, "w") as f:
            json.dump(data, f, indent=2, sort_keys=True)

    # ---- decompress-on-first-forward hook -------------------------------------------------------
    def add_decompress_hook(self, model: torch.nn.Module):
        def ct_decompress_hook(model, args):
            self.decompress_model(model, distributed=False)   # never a collective: only some ranks may run forward

        model.ct_decompress_hook = model.register_forward_pre_hook(ct_decompress_hook)

    def remove_decompression_hook(self, model: torch.nn.Module):
        if hasattr(model, "ct_decompress_hook"):
            model.ct_decompress_hook.remove()
            delattr(model, "ct_decompress_hook")
