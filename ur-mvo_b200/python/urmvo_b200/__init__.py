"""urmvo_b200 — Python (ctypes) binding of the C ABI in include/urmvo_b200.h.

Used by tests/, bench.py and __graft_entry__.py.  The product is the shared library
ur-mvo_b200/lib/liburmvo_b200.so (hand-written sm_100a CUDA); this module only marshals numpy
arrays into it.  There is no CPU fallback: importing works without a GPU (so that symbol checks can
run anywhere), but creating a Context without a usable B200 raises.
"""
from .capi import (Context, BAPlan, PosePlan, TVPlan, ShardedBAPlan, BAStats, TVStats, BAOptions, UrmvoError,
                   lib_path, load_library, build_library, EXPORTED_SYMBOLS, nccl_unique_id, ba_covisibility,
                   shard_points, pack_ba_batch, FMPlan, FMStats, pack_fm_batch, DeviceMap, PnPStats)

__all__ = ["Context", "BAPlan", "PosePlan", "TVPlan", "ShardedBAPlan", "nccl_unique_id", "ba_covisibility",
           "shard_points", "pack_ba_batch", "FMPlan", "FMStats", "pack_fm_batch", "BAStats", "TVStats", "BAOptions", "UrmvoError",
           "lib_path", "load_library", "build_library", "EXPORTED_SYMBOLS", "DeviceMap", "PnPStats"]
