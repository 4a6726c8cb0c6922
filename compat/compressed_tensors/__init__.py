"""
Drop-in alias: with `<repo>/compat` on sys.path, `import compressed_tensors` (and every `compressed_tensors.<sub.module>` path of
SURVEY.md Appendix C) resolves to the B200 engine, so code written against vllm-project/compressed-tensors -- including the
reference's own hot-path tests -- runs on this package without editing its imports:

    PYTHONPATH=/path/to/repo/compat:/path/to/repo python -m pytest <reference>/tests/test_compressors/test_pack_quant.py

The package's modules are loaded under the `compressed_tensors.*` names (its `__path__` points at compressed_tensors_b200/), and all
imports inside the package are relative, so one consistent namespace results (its own registry, one native library handle).
Do not mix it with `import compressed_tensors_b200` in the same process: that would be a second copy of the registry.
"""
import importlib
import os

_IMPL = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "compressed_tensors_b200")
__path__ = [_IMPL]
__version__ = "0.1.0"

_LAZY = {
    "ModelCompressor": "compressors", "BaseCompressor": "compressors", "PackedQuantizationCompressor": "compressors",
    "IntQuantizationCompressor": "compressors", "FloatQuantizationCompressor": "compressors", "NaiveQuantizationCompressor": "compressors",
    "NVFP4PackedCompressor": "compressors", "MXFP4PackedCompressor": "compressors", "MXFP8QuantizationCompressor": "compressors",
    "compress_module": "compressors", "decompress_module": "compressors",
    "CompressionFormat": "config", "SparsityCompressionConfig": "config", "SparsityStructure": "config",
    "QuantizationArgs": "quantization", "QuantizationScheme": "quantization", "QuantizationConfig": "quantization",
    "QuantizationStatus": "quantization", "QuantizationStrategy": "quantization", "QuantizationType": "quantization",
}


_STAR = ("compressors", "config", "quantization", "registry", "utils")   # the sub-packages the reference star-exports (__init__.py:6-22)


def __getattr__(name):
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
