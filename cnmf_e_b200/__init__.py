"""cnmf_e_b200 -- B200-native CNMF-E alternating-update hot path (ring background -> spatial -> temporal + OASIS).

Python host-side mirror of the reference's MATLAB interface for this path; all compute is in
libcnmfe_b200.so (hand-written CUDA for sm_100a, C ABI in include/cnmfe_b200.h)."""
from ._lib import CnmfeError, LIB_PATH  # noqa: F401
