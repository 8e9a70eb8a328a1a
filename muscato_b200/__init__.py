"""muscato_b200 -- B200-native screen -> group -> confirm -> combine hot path of Muscato.

The product is libmuscato_b200.so (hand-written sm_100a CUDA behind the C ABI in
include/muscato_b200.h).  This package is the thin host-side mirror of the reference's
config / file contracts used by the tests, the benchmark and the stage tools."""
from .config import Config  # noqa: F401

__all__ = ["Config"]
