"""
compressed_tensors_b200 -- B200-native (sm_100a) engine behind the compressed-tensors
compress()/decompress() and quantize()/dequantize() path.

Layout
  csrc/            hand-written CUDA kernels + the C ABI (include/ct_b200.h) -> libct_b200.so
  _native.py       ctypes binding of the C ABI (fails loudly when the library / GPU is missing)
  ops.py           tensor-level front end mirroring the reference's per-tensor functions
  quantization/, compressors/, config/, registry/, distributed/, utils/
                   host-side mirror of the reference's plugin / operator interface for this path
"""
__version__ = "0.1.0"


# The reference star-exports its sub-packages at the top level (src/compressed_tensors/__init__.py:6-22).  Here the same names
# resolve lazily, so that `import compressed_tensors_b200` stays cheap and does not need torch / the native library.
_LAZY = {
    "ModelCompressor": "compressors", "BaseCompressor": "compressors", "PackedQuantizationCompressor": "compressors",
    "IntQuantizationCompressor": "compressors", "FloatQuantizationCompressor": "compressors", "NaiveQuantizationCompressor": "compressors",
    "NVFP4PackedCompressor": "compressors", "MXFP4PackedCompressor": "compressors", "MXFP8QuantizationCompressor": "compressors",
    "compress_module": "compressors", "decompress_module": "compressors",
    "CompressionFormat": "config", "SparsityCompressionConfig": "config", "SparsityStructure": "config",
    "QuantizationArgs": "quantization", "QuantizationScheme": "quantization", "QuantizationConfig": "quantization",
    "QuantizationStatus": "quantization", "QuantizationStrategy": "quantization", "QuantizationType": "quantization",
}


_STAR = ("compressors", "config", "quantization", "registry", "utils")   # the sub-packages the reference star-exports


def __getattr__(name):
    import importlib

    if name in _LAZY:
        return getattr(importlib.import_module(f"{__name__}.{_LAZY[name]}"), name)
    if not name.startswith("_"):
        for sub in _STAR:
            try:
                mod = importlib.import_module(f"{__name__}.{sub}")
            except ImportError:      # hasattr() on the package must stay a question, not an import failure
                continue
            if name in getattr(mod, "__all__", ()) or (not hasattr(mod, "__all__") and hasattr(mod, name)):
                return getattr(mod, name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
