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
