"""vp8b200 - Python host-side helpers around the C ABI of libvp8b200.so.

PyTorch is not needed by the product path; this package is ctypes + numpy:
  abi      ctypes binding of include/vp8b200.h (fails loudly if the library is missing)
  recfile  reader/writer for .rec record dumps (include/vp8b200_recfile.h)
  frames   YV12 buffer geometry, visible-area crop and the MD5 that `vpxdec --md5` prints
"""
from . import recfile, frames  # noqa: F401
