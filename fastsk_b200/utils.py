"""FASTA-like ingest: the host-side mirror of the reference's ``fastsk.utils``.

Same public names and behaviour as /root/reference/src/fastsk/utils.py:
  * ``Vocabulary``   (utils.py:5-36)  -- token -> id, ids handed out in first-seen order
                                         starting at 1; id 0 is reserved (utils.py:13).
  * ``FastaUtility`` (utils.py:39-104) -- ``read_data(path)`` -> (X, Y) with alternating
                                         ``>label`` / sequence lines, lower-cased,
                                         ``shortest_seq(path)``.
plus ``read_encoded``: the same parse straight into the flat
``codes`` / ``offsets`` arrays the C ABI takes, skipping the list-of-lists of Python ints.
"""
from __future__ import annotations

import numpy as np


class Vocabulary(object):
    """Maps tokens to integer ids (first-seen order, 0 reserved for "unknown")."""

    def __init__(self):
        self._token2idx = {0: 0}
        self._size = 1

    def add(self, token):
        """Return the id of ``token``, assigning the next free id on first sight."""
        idx = self._token2idx.get(token)
        if idx is None:
            idx = self._size
            self._token2idx[token] = idx
            self._size += 1
        return idx

    def size(self):
        return self._size

    def __str__(self):
        return str(self._token2idx)


class FastaUtility:
    def __init__(self, vocab=None):
        self._vocab = Vocabulary() if vocab is None else vocab

    @staticmethod
    def _records(data_file, regression):
        """Yield (label, lower-cased sequence string); label/sequence lines alternate."""
        label = None
        expect_label = True
        with open(data_file, "r") as f:
            for line in f:
                line = line.strip().lower()
                if expect_label:
                    parts = line.split(">")
                    assert len(parts) == 2
                    if regression:
                        label = parts[1]
                    else:
                        label = int(parts[1])
                        assert label in [-1, 0, 1]
                    expect_label = False
                else:
                    yield label, line
                    expect_label = True

    def read_data(self, data_file, vocab="inferred", regression=False):
        """Return (X, Y): X = list of id lists, Y = list of labels (reference utils.py:50-96)."""
        assert vocab.lower() in ["dna", "protein", "inferred"]
        X, Y = [], []
        add = self._vocab.add
        for label, seq in self._records(data_file, regression):
            Y.append(label)
            X.append([add(ch) for ch in seq])
        assert len(X) == len(Y)
        return X, Y

    def read_encoded(self, data_file, regression=False):
        """Return (codes int32[sum len], offsets int64[n+1], labels) with the same ids as read_data."""
        chunks, labels, lens = [], [], []
        lut = np.full(256, -1, dtype=np.int32)        # code point -> id, grown on demand
        for tok, idx in self._vocab._token2idx.items():
            if isinstance(tok, str) and len(tok) == 1:
                if ord(tok) >= len(lut):
                    lut = np.concatenate([lut, np.full(ord(tok) + 1 - len(lut), -1, dtype=np.int32)])
                lut[ord(tok)] = idx
        for label, seq in self._records(data_file, regression):
            labels.append(label)
            raw = np.frombuffer(seq.encode("utf-32-le"), dtype=np.uint32)
            if len(raw) and raw.max() >= len(lut):
                lut = np.concatenate([lut, np.full(int(raw.max()) + 1 - len(lut), -1, dtype=np.int32)])
            ids = lut[raw]
            if (ids < 0).any():                       # new characters: ids in first-seen order
                uniq, first = np.unique(raw[ids < 0], return_index=True)
                for cp in uniq[np.argsort(first)]:
                    lut[cp] = self._vocab.add(chr(int(cp)))
                ids = lut[raw]
            chunks.append(ids)
            lens.append(len(raw))
        offsets = np.zeros(len(lens) + 1, dtype=np.int64)
        np.cumsum(np.asarray(lens, dtype=np.int64), out=offsets[1:])
        codes = np.concatenate(chunks).astype(np.int32) if chunks else np.zeros(0, dtype=np.int32)
        return codes, offsets, labels

    def shortest_seq(self, data_file):
        X, _ = self.read_data(data_file)
        return min(len(x) for x in X)
