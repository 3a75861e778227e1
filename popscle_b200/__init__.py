"""popscle_b200 — B200-native demuxlet / freemuxlet genotype-likelihood engine.

Product code only: the CUDA library (csrc/ -> libpopscle_b200.so, C ABI in include/popscle_b200.h),
its ctypes mirror (capi), the synthetic pileup generator (synth), the multi-GPU drivers (dist) and
the C++ CLI host (host/).  The CPU oracle lives in /oracle and is never imported from here.
"""
from .capi import (Context, Multi, bind_to_device, DevicePileup, Pileup, PsclError, RawGeno, DEMUX_CELL_DTYPE, FMX_CELL_DTYPE, TYPE_NAMES,
                   load_library, LIB_PATH)

__all__ = ["Context", "Multi", "bind_to_device", "DevicePileup", "Pileup", "PsclError", "RawGeno", "DEMUX_CELL_DTYPE", "FMX_CELL_DTYPE",
           "TYPE_NAMES", "load_library", "LIB_PATH"]
