"""pantas_b200 -- B200-native `pantas augment`.

One hot path and nothing else: the GAF loop of
/root/reference/scripts/alignments_augmentation_from_gaf.py runs as hand-written
sm_100a CUDA kernels (pantas_b200/csrc) behind the C ABI of
include/pantas_aug.h; the CLI, GFA loading and GFA writing stay in Python.
"""
from .errors import NativeLibraryError, PantasDataError, PantasError, UnsupportedInput  # noqa: F401

__all__ = ["PantasError", "PantasDataError", "UnsupportedInput", "NativeLibraryError"]
__version__ = "0.1.0"
