"""pure_zlib_b200 -- B200-native zlib inflate behind pure-zlib's `Codec.Compression.Zlib` API.

The product is `libpzcuda.so` (hand-written sm_100a kernels + C ABI, see include/pzcuda.h).
This package is the host-side mirror of the reference's module interface
(`decompress`, `decompressIncremental`, `ZlibDecoder`, `DecompressionError`) bound to that
ABI with ctypes, standing in for the Haskell shim (haskell/) where no GHC exists.

There is no CPU decode path: every call goes to the GPU or raises.
"""
from .zlib import (ChecksumError, Chunk, DecompError, DecompressionError, DecompressionError_, Done, FormatError,  # noqa: F401
                   HeaderError, HuffmanTreeError, Left, NeedMore, ReferenceBottom, Right, compute_code_values,
                   decompress, decompress_batch, decompress_incremental, decompress_many, IncrementalSet,
                   decompress_gzip, decompress_raw, ZLIB, GZIP, RAW)

__all__ = ["decompress", "decompress_incremental", "decompress_batch", "DecompressionError", "HuffmanTreeError",
           "FormatError", "DecompressionError_", "HeaderError", "ChecksumError", "ReferenceBottom", "NeedMore", "Chunk",
           "Done", "DecompError", "Left", "Right", "compute_code_values", "decompress_many", "IncrementalSet",
           "decompress_gzip", "decompress_raw", "ZLIB", "GZIP", "RAW"]
