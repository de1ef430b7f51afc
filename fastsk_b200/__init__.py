"""fastsk_b200: B200-native gapped k-mer kernel-matrix build behind FastSK's Python API.

    from fastsk_b200 import FastSK, FastaUtility

mirrors ``from fastsk import FastSK, FastaUtility`` (reference src/fastsk/__init__.py:1-2).
"""
from .fastsk import FastSK, pinned_empty, shard_rows, shared_output
from .utils import FastaUtility, Vocabulary

__version__ = "0.2"
__all__ = ["FastSK", "FastaUtility", "Vocabulary", "pinned_empty", "shared_output", "shard_rows"]
