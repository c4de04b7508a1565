"""bronko_b200 — B200 (sm_100a) implementation of bronko's k-mer→pileup path.

The product is bronko_b200/csrc (CUDA kernels + C ABI, include/bronko_b200.h); this package is the
Python mirror of the reference's seam on top of that ABI.  There is no CPU path: importing works
anywhere, creating a context requires the compiled library and a B200.
"""
from .api import Bronko, CallArgs, DecodedReads, Sample, clean_sample_id, pack_reads  # noqa: F401
from ._lib import BkError  # noqa: F401

__version__ = "0.1.0"
