"""
Typed aliases of the reference's `utils/type.py`: `TorchDtype` (a pydantic field type for torch.dtype, written as "torch.float16" style
strings in config.json) and `TensorStateDict`.  The field type lives next to its main user, QuantizationArgs
(quantization/quant_args.py); this module is the reference's import path for it.
"""
from ..quantization.quant_args import TorchDtype
from .helpers import TensorStateDict

__all__ = ["TorchDtype", "TensorStateDict"]
