"""Row index -> paragraph id, array-based.

The reference turns the int64 result ``I`` of ``index.search`` into paragraph ids by loading a JSON dict keyed by
``str(row)`` and walking ``I`` in two Python loops (retrieval/eval_retrieval.py:68-76; the dict is written by
retrieval/gen_index_id_map.py:3-9 in corpus order, so key ``str(j)`` is simply row j).  For 21M rows that dict costs
minutes and gigabytes; a flat array indexed by row gives the same lists.  Host-side only — nothing here touches the GPU —
and the result is the reference's, element for element, including its failure mode: a padding id of -1 (k > ntotal) is a
``KeyError`` there and here (never a silent wrap-around to the last row).
"""
from __future__ import annotations

import json

import numpy as np


class IdMap:
    """``ids[j]`` is the paragraph id of corpus row ``j`` (any JSON scalar: str or int)."""

    def __init__(self, ids):
        self.ids = np.asarray(ids, dtype=object)
        assert self.ids.ndim == 1

    def __len__(self):
        return int(self.ids.shape[0])

    @classmethod
    def from_json(cls, path):
        """Read the reference's ``idx_id.json`` (``{"0": id0, "1": id1, ...}``, gen_index_id_map.py:7-9)."""
        m = json.load(open(path))
        n = len(m)
        ids = np.empty(n, dtype=object)
        try:
            for j in range(n):
                ids[j] = m[str(j)]
        except KeyError as e:
            raise ValueError(f"{path}: keys are not the dense range 0..{n - 1} (missing {e})") from None
        return cls(ids)

    @classmethod
    def from_jsonl(cls, path, field="id"):
        """Straight from the corpus file gen_index_id_map.py reads (one JSON object per line, row = line number)."""
        with open(path) as f:
            return cls([json.loads(line)[field] for line in f])

    def save(self, path):
        np.save(path, self.ids, allow_pickle=True)

    @classmethod
    def load(cls, path):
        return cls(np.load(path, allow_pickle=True))

    def convert(self, I):
        """``convert_idx2id(I)`` of eval_retrieval.py:68-76: list (per query) of lists (per rank) of paragraph ids."""
        I = np.asarray(I)
        if I.size and (I.min() < 0 or I.max() >= len(self)):
            bad = I[(I < 0) | (I >= len(self))].ravel()[0]
            raise KeyError(str(int(bad)))
        return self.ids[I].tolist() if I.size else [[] for _ in range(I.shape[0])]


def convert_idx2id(idxs, mapping_path="../pretrained_models/idx_id.json"):
    """Same signature and default path as the reference's function."""
    return IdMap.from_json(mapping_path).convert(idxs)
