"""module path of the reference (quantization/utils/fp4_utils.py:77-98): round to the nearest E2M1 value, same dtype; one CUDA kernel
(`ct_cast_to_fp4`, include/ct_b200.h)"""
from ...ops import cast_to_fp4  # noqa: F401

__all__ = ["cast_to_fp4"]
