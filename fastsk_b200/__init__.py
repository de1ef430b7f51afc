"""fastsk_b200: B200-native gapped k-mer kernel-matrix build behind FastSK's Python API.

    from fastsk_b200 import FastSK, FastaUtility

mirrors ``from fastsk import FastSK, FastaUtility`` (reference src/fastsk/__init__.py:1-2).
"""
from .fastsk import FastSK
from .utils import FastaUtility, Vocabulary

__version__ = "0.1"
__all__ = ["FastSK", "FastaUtility", "Vocabulary"]
