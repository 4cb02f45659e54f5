"""rt3-b200: B200-native render core behind RayTracer-3's Renderer interface.

The product is the CUDA library ``csrc/librt3cuda.so`` (C ABI in
``include/rt3cuda.h``) and the C++ host backend in ``host/``; this Python
package is plumbing for tests and benchmarks (ctypes bindings, synthetic
scenes). It is loaded through the top-level ``rt3_b200`` module because the
directory name is not a valid Python identifier.
"""
from . import abi, distributed, scenes  # noqa: F401
