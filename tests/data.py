"""Seeded synthetic inputs shared by the tests (shapes from SURVEY.md §8d)."""
import numpy as np

D = 128


def corpus(n, seed=1234, kind="normal"):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, D), dtype=np.float32)
    if kind == "fp16":  # embeddings produced with --fp16 (get_embed.py:147-151): exactly representable in half
        x = x.astype(np.float16).astype(np.float32)
    elif kind == "unit":
        x /= np.linalg.norm(x, axis=1, keepdims=True)
    elif kind == "skewed":  # norms spread over two orders of magnitude
        x *= np.exp(rng.uniform(-2.3, 2.3, size=(n, 1))).astype(np.float32)
    return np.ascontiguousarray(x, dtype=np.float32)


def queries(n, seed=4321, kind="normal"):
    return corpus(n, seed=seed, kind=kind)
