"""pysgs -- thin Python glue over libsgs_cuda.so (ctypes) for tests and bench.

The product is the C-ABI library (include/sgs_abi.h) and its C++ host mirror; this package only
binds it for the Python-side harness.  Importing `pysgs.binding` fails loudly when the CUDA library
has not been built -- there is no CPU fallback.
"""
