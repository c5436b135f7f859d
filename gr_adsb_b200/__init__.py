"""gr-adsb_b200: the ADS-B burst-detect + PPM-demod path of gr-adsb, on a B200.

`framer` and `demod` mirror `gnuradio.adsb.framer` / `gnuradio.adsb.demod`
(python/adsb/__init__.py:24-25); `burst` is the two fused into one block (optionally taking complex64 IQ); `Context` is the raw ctypes handle onto `include/adsb_b200.h`.
"""
from ._native import Context, AdsbError, HIT_DTYPE, unpack_bits, build, lib  # noqa: F401
from ._native import HIT_DEMOD_VALID, HIT_CRC_OK, HIT_LATE_TAG, HIT_FEC_FIXED, CARRY_USED_ON_DEVICE  # noqa: F401
from ._native import FEC_NONE, FEC_CONSERVATIVE, FMT_F32, FMT_C64, FMT_SC16, FMT_CU8  # noqa: F401
from ._native import (SHARD_PH_SIGNAL, SHARD_PH_SCAN, SHARD_PH_EXPORT, SHARD_PH_PLAN, SHARD_PH_CHECK,  # noqa: F401
                      SHARD_PH_ALL, E_PEER)


def __getattr__(name):
    # blocks import lazily: they need the native library, `synth` and the ABI constants do not
    if name in ("framer", "demod", "burst", "pdu_unbatch", "split_batch", "feed_decoder"):
        from . import blocks
        return getattr(blocks, name)
    raise AttributeError(name)
