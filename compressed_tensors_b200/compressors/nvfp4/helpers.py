"""fp4 nibble packing (mirror of compressors/nvfp4/helpers.py:108-193); the CUDA library does the work."""
from ...ops import pack_fp4_to_uint8, unpack_fp4_from_uint8

__all__ = ["pack_fp4_to_uint8", "unpack_fp4_from_uint8"]
