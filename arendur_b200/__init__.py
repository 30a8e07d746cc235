"""arendur_b200 — B200-native path-tracing core behind arendur's renderer API.

The product is the CUDA library arendur_b200/libarn_b200.so (C-ABI: include/arn.h,
include/arn_host.h; sources in arendur_b200/csrc).  This Python package is the test / bench
harness over that ABI; importing `arendur_b200.api` fails loudly when the library is not built.
"""
__version__ = "0.1"
