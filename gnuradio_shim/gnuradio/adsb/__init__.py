"""`gnuradio.adsb` with the burst path on a B200: same names as gr-adsb's python/adsb/__init__.py:24-26."""
from gr_adsb_b200 import framer, demod, burst, pdu_unbatch      # noqa: F401

try:                                                             # the unchanged Python decoder, wherever gr-adsb put it
    from adsb_decoder import decoder                             # noqa: F401  (a copy placed on the path by the user)
except Exception:                                                # pragma: no cover
    try:
        from decoder import decoder                              # noqa: F401
    except Exception:
        decoder = None
