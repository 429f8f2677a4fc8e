"""ORACLE stand-in so the UNMODIFIED reference (when /root/reference is mounted) can be
imported on top of the oracle arithmetic:
  PYTHONPATH=oracle/standin:/root/reference/curdleproofs:/root/reference/merlin_transcripts
Used only by tests and by oracle/gen_golden.py."""
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _root not in sys.path:
    sys.path.insert(0, _root)
from oracle.ark_surface import G1Point, Scalar  # noqa: E402,F401
