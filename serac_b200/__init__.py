"""serac_b200: B200-native (sm_100a) evaluation engine behind serac::Functional.

The product is libserac_b200.so (C ABI in include/serac_b200.h, CUDA kernels in serac_b200/include + serac_b200/csrc);
`serac_b200.capi` is its ctypes binding.  Importing the package does not need a GPU; calling it does."""
from . import capi  # noqa: F401
