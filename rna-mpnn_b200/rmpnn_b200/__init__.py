"""rmpnn_b200 -- host-side binding of the B200-native RNA-MPNN forward path.

The compute lives in ``librmpnn_b200.so`` (hand-written sm_100a CUDA behind the C ABI
declared in ``include/rmpnn_b200.h``).  This package only marshals torch tensors
(device memory, streams) into that ABI.  There is no CPU or PyTorch fallback: every
op raises if the library is missing or a tensor is not on a CUDA device.
"""
from ._lib import lib, lib_path, LibraryMissing  # noqa: F401
