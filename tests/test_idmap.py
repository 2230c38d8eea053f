"""CPU: proqa_b200.idmap gives the lists the reference's convert_idx2id gives (eval_retrieval.py:68-76), restated here as the
two-loop dict walk, on the committed eval fixture and on edge cases."""
import json

import numpy as np
import pytest

from proqa_b200.idmap import IdMap, convert_idx2id
from tests.golden_util import load_eval_fixture


def dict_walk(idxs, mapping):                    # the reference's algorithm
    return [[mapping[str(j)] for j in row] for row in idxs]


def test_same_lists_as_the_dict_walk(tmp_path):
    rng = np.random.default_rng(0)
    n = 5000
    mapping = {str(j): (f"doc{j * 7919 % n}_p{j}" if j % 3 else j * 11) for j in range(n)}    # str and int ids both occur upstream
    p = tmp_path / "idx_id.json"
    json.dump(mapping, open(p, "w"))
    I = rng.integers(0, n, size=(37, 80)).astype(np.int64)
    m = IdMap.from_json(p)
    assert len(m) == n
    assert m.convert(I) == dict_walk(I, mapping)
    assert convert_idx2id(I, str(p)) == dict_walk(I, mapping)
    m.save(tmp_path / "ids.npy")
    assert IdMap.load(tmp_path / "ids.npy").convert(I) == dict_walk(I, mapping)


def test_padding_id_is_a_key_error_like_upstream(tmp_path):
    m = IdMap([f"d{j}" for j in range(10)])
    I = np.array([[3, 2, -1]], np.int64)
    with pytest.raises(KeyError):
        m.convert(I)
    with pytest.raises(KeyError):
        dict_walk(I, {str(j): f"d{j}" for j in range(10)})
    with pytest.raises(KeyError):
        m.convert(np.array([[10]], np.int64))
    assert m.convert(np.empty((2, 0), np.int64)) == [[], []]


def test_sparse_keys_are_rejected(tmp_path):
    p = tmp_path / "m.json"
    json.dump({"0": "a", "2": "b"}, open(p, "w"))
    with pytest.raises(ValueError):
        IdMap.from_json(p)


def test_jsonl_corpus_and_eval_fixture(tmp_path):
    fx = load_eval_fixture()
    n = fx["xb"].shape[0]
    p = tmp_path / "para_doc.db"
    with open(p, "w") as f:
        for j in range(n):
            f.write(json.dumps({"id": f"d{j}", "text": "x"}) + "\n")
    m = IdMap.from_jsonl(p)
    assert m.convert(fx["I"]) == dict_walk(fx["I"], {str(j): f"d{j}" for j in range(n)})
