"""tensorium_b200 — B200 (sm_100a) native backend for the convolution hot path of Tensorium.

The product is libtensorium_b200.so (C ABI: include/tensorium_b200.h); this package holds its CUDA
sources (csrc/), the in-tree build (build.py) and the host-side mirror of the reference's operator
interface (nncuda.TNNCuda, tensor.TSingleTensor, layers.*).  Nothing here imports oracle/.
"""
from .nncuda import TNNCuda, CudaError  # noqa: F401
