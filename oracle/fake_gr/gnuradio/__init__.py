"""Stand-in for the `gnuradio` package — TEST INFRASTRUCTURE ONLY.

GNU Radio is not installed in the build container.  The unmodified reference
blocks (`/root/reference/python/adsb/{framer,demod,decoder}.py`) only touch the
handful of `gr.sync_block` / `pmt` entry points listed in SURVEY.md Appendix B;
this package provides exactly those so `oracle/run_reference.py` can execute the
reference source as-is to generate the golden vectors under `tests/golden/`.
Nothing in the product package imports this.
"""
from . import gr  # noqa: F401
