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


# ASCII bytes str.strip() removes (the reference strips every line, utils.py:50-96): a file holding any of them inside its
# sequence lines goes to the record parser
_STRIPPED = (9, 11, 12, 13, 28, 29, 30, 31, 32)


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
        """Return (codes int32[sum len], offsets int64[n+1], labels) with the same ids as read_data.

        Plain ASCII files without blanks or carriage returns (every bundled data set) are parsed in one vectorised
        pass over the file's bytes; anything else goes record by record."""
        fast = self._read_encoded_bytes(data_file, regression)
        if fast is not None:
            return fast
        return self._read_encoded_records(data_file, regression)

    def _read_encoded_bytes(self, data_file, regression):
        raw = np.fromfile(data_file, dtype=np.uint8)
        if raw.size == 0:
            return np.zeros(0, dtype=np.int32), np.zeros(1, dtype=np.int64), []
        ends = np.flatnonzero(raw == 10)
        if raw[-1] != 10:
            ends = np.append(ends, raw.size)
        starts = np.concatenate(([0], ends[:-1] + 1))
        if len(starts) % 2:
            return None                                   # a label without a sequence: let the record parser complain
        ls, le = starts[0::2], ends[0::2]                 # label lines
        ss, se = starts[1::2], ends[1::2]                 # sequence lines
        llen = le - ls
        if (llen < 2).any() or (raw[ls] != ord(">")).any() or llen.max() > 32:
            return None
        if regression:
            labels = [raw[a + 1:b].tobytes().decode().strip().lower() for a, b in zip(ls, le)]
            if any(">" in lab for lab in labels):
                return None
        else:
            lab = np.full(len(ls), 2, dtype=np.int64)     # 2 = not one of the three short forms
            b1 = raw[ls + 1]
            b2 = raw[np.minimum(ls + 2, raw.size - 1)]
            short = (llen == 2) & ((b1 == ord("0")) | (b1 == ord("1")))
            lab[short] = b1[short].astype(np.int64) - 48
            lab[(llen == 3) & (b1 == ord("-")) & (b2 == ord("1"))] = -1
            for i in np.flatnonzero(lab == 2):             # "+1", "01", ... : the reference's int()
                text = raw[ls[i] + 1:le[i]].tobytes().decode()
                if ">" in text:
                    return None
                lab[i] = int(text)
                assert lab[i] in (-1, 0, 1)
            labels = lab.tolist()
        lens = (se - ss).astype(np.int64)
        offsets = np.zeros(len(lens) + 1, dtype=np.int64)
        np.cumsum(lens, out=offsets[1:])
        # bytes of all sequence lines, in file order: everything but the newlines and the (short) label lines
        keep = np.ones(raw.size, dtype=bool)
        keep[ends[ends < raw.size]] = False
        for j in range(int(llen.max())):
            at = ls + j
            keep[at[at < le]] = False
        seq = raw[keep]
        assert seq.size == int(offsets[-1])
        # one table: byte -> id of its lower-cased character; -2 = a byte the per-line strip() / decoding would have to deal
        # with (blanks, carriage returns, non-ASCII: leave those files to the record parser); -1 = not in the vocabulary yet
        lower = np.arange(256)
        lower[65:91] += 32
        table = np.full(256, -1, dtype=np.int32)
        for tok, k in self._vocab._token2idx.items():
            if isinstance(tok, str) and len(tok) == 1 and ord(tok) < 128:
                table[ord(tok)] = k
        known = table.copy()
        uniq, first = np.unique(lower[seq[:65536]], return_index=True)     # first-seen order: the head settles almost all
        pending = [int(v) for v in uniq[np.argsort(first)] if v < 128 and v not in _STRIPPED and known[v] < 0]
        next_id = self._vocab.size()
        for v in pending:
            known[v] = next_id
            next_id += 1
        table = known[lower]
        table[128:] = -2
        table[list(_STRIPPED)] = -2
        codes = table[seq]
        lo = int(codes.min()) if codes.size else 0
        if lo == -2:
            return None
        if lo == -1:                                       # characters that first appear after the head, in order of appearance
            late = np.flatnonzero(codes == -1)
            vals, first = np.unique(lower[seq[late]], return_index=True)
            for v in vals[np.argsort(first)]:
                pending.append(int(v))
                known[int(v)] = next_id
                next_id += 1
            codes = known[lower][seq]
        for v in pending:                                  # commit to the vocabulary only now that the file is accepted
            self._vocab.add(chr(v))
        return codes, offsets, labels

    def _read_encoded_records(self, data_file, regression=False):
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
