#!/usr/bin/env python
"""Generate the committed golden vectors.  Run HERE (build container; needs /root/reference), never on the GPU box.

    python tests/golden/make_fixtures.py

1. eval_fixture.npz — the reference's own script /root/reference/retrieval/eval_retrieval.py is run UNMODIFIED
   (as a subprocess, cwd = a scratch tree that has ../pretrained_models/idx_id.json, exactly as the script expects)
   on synthetic inputs with ``import faiss`` resolved to the FAISS-1.6.3 restatement (tests/golden/_oracle_faiss).
   Stored: the inputs (float16: exactly what get_embed.py --fp16 would have written), the I the script obtained, the
   recall lines it printed, and — from the reference's own para_has_answer/SimpleTokenizer/DocDB — the
   [question, paragraph] "has answer" bit matrix, so the GPU box can re-derive the recall lines from its own I with
   pure integer work and no reference code.
2. known_answers.json — hand-checkable exact cases (small integers: every product and sum is exact in fp32, so any
   correct implementation, any accumulation order, must return these bits).
3. trec_fixture.npz — retrieval/trec_process.py: retrieve_topk() (the k = 10000 call site, trec_process.py:69-94) imported and
   called UNMODIFIED on synthetic inputs with the same ``faiss`` binding.  Stored: the seed of the inputs (regenerated and
   checksummed by the tests), the labels, the I the function obtained (from the file it wrote) and the line it printed.
"""
import json
import os
import sqlite3
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/retrieval"
sys.path.insert(0, ROOT)

N, NQ, K, D = 2000, 24, 80, 128
MIN_GAP = 6e-5  # absolute; scores are O(20), fp32 dot-product error is O(2e-6): 30x head-room
WORDS = ["alder", "birch", "cedar", "dogwood", "elm", "fir", "ginkgo", "hazel", "ironwood", "juniper", "kapok", "larch",
         "maple", "nutmeg", "oak", "pine", "quince", "rowan", "spruce", "teak", "upas", "viburnum", "willow", "yew", "zelkova"]


def make_inputs(seed):
    rng = np.random.default_rng(seed)
    xb = rng.standard_normal((N, D)).astype(np.float16)
    xq = rng.standard_normal((NQ, D)).astype(np.float16)
    # plant structure: query i is close to rows planted[i]; so recall is neither 0 nor 1 everywhere
    for i in range(NQ):
        for j in rng.choice(N, size=3, replace=False):
            xb[j] = (0.6 * xq[i].astype(np.float32) + 0.8 * xb[j].astype(np.float32)).astype(np.float16)
    return xb, xq


def unambiguous(xb, xq):
    """Every rank boundary of the fp64 top-(K+1) is much wider than fp32 rounding noise, so I does not depend on the
    accumulation order of whoever computes the fp32 scores (OpenBLAS here, the sequential chain on the GPU): the
    committed I is THE answer, with no near-tie to argue about."""
    S = xq.astype(np.float64) @ xb.astype(np.float64).T
    top = -np.sort(-S, axis=1)[:, :K + 1]
    gap = top[:, :-1] - top[:, 1:]
    return bool((gap > MIN_GAP).all())


def find_inputs():
    seed = 20240
    while True:
        xb, xq = make_inputs(seed)
        if unambiguous(xb, xq):
            return seed, xb, xq
        seed += 1


def build_eval_tree(tmp, seed, xb, xq):
    """The scratch tree retrieval/eval_retrieval.py expects (cwd = tmp/retrieval): paragraphs in sqlite, the idx->id JSON of
    gen_index_id_map.py, the question file, both embedding files.  Deterministic in `seed`.  -> (db, qa, answers, texts)"""
    rng = np.random.default_rng(seed + 7)
    os.makedirs(os.path.join(tmp, "retrieval"), exist_ok=True)
    os.makedirs(os.path.join(tmp, "pretrained_models"), exist_ok=True)
    # paragraphs: random word salads; the answer string of question i is planted in a few paragraphs
    answers = [[f"{WORDS[i % len(WORDS)]} {WORDS[(7 * i + 3) % len(WORDS)]} {i}"] for i in range(NQ)]
    texts = []
    for j in range(N):
        texts.append(" ".join(rng.choice(WORDS, size=12)) + f" . Paragraph {j} .")
    S = xq.astype(np.float64) @ xb.astype(np.float64).T
    order = np.argsort(-S, axis=1)
    for i in range(NQ):
        r = int(rng.integers(0, 140))  # answer sits at true rank r (beyond K for some questions: those are misses)
        j = int(order[i, r])
        texts[j] += f" The Answer Is {answers[i][0].upper()} , indeed ."
    ids = [f"doc_{j:05d}" for j in range(N)]
    db = os.path.join(tmp, "paras.db")
    if os.path.exists(db):
        os.remove(db)
    con = sqlite3.connect(db)
    con.execute("CREATE TABLE documents (id PRIMARY KEY, text)")
    con.executemany("INSERT INTO documents VALUES (?,?)", list(zip(ids, texts)))
    con.commit()
    con.close()
    # gen_index_id_map.py:3-9 semantics: {line number: sample['id']} dumped as JSON (int keys become strings)
    json.dump({j: ids[j] for j in range(N)}, open(os.path.join(tmp, "pretrained_models", "idx_id.json"), "w"))
    qa = os.path.join(tmp, "qa.jsonl")
    with open(qa, "w") as f:
        for i in range(NQ):
            f.write(json.dumps({"question": f"which tree number {i} ?", "answer": answers[i]}) + "\n")
    np.save(os.path.join(tmp, "para_embed.npy"), xb)
    np.save(os.path.join(tmp, "query_embed.npy"), xq)
    return db, qa, answers, texts


def main():
    seed, xb, xq = find_inputs()
    tmp = tempfile.mkdtemp(prefix="proqa_golden_")
    db, qa, answers, texts = build_eval_tree(tmp, seed, xb, xq)

    env = dict(os.environ)
    env["PYTHONPATH"] = os.path.join(HERE, "_oracle_faiss") + os.pathsep + env.get("PYTHONPATH", "")
    env["PROQA_GOLDEN_DUMP"] = os.path.join(tmp, "dump.npz")
    out = subprocess.run([sys.executable, os.path.join(REF, "eval_retrieval.py"), qa, os.path.join(tmp, "para_embed.npy"),
                          os.path.join(tmp, "query_embed.npy"), db, "--topk", str(K), "--num-workers", "2"],
                         cwd=os.path.join(tmp, "retrieval"), env=env, capture_output=True, text=True, check=True)
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("Top ")]
    assert len(lines) == 5, out.stdout + out.stderr
    dump = np.load(os.path.join(tmp, "dump.npz"))

    # has-answer matrix with the reference's own scorer (eval_retrieval.py:27-45)
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(HERE, "_oracle_faiss"))
    import importlib
    er = importlib.import_module("eval_retrieval")
    er.init(db)
    has = np.zeros((NQ, N), dtype=bool)
    for i in range(NQ):
        for j in range(N):
            if answers[i][0].split()[0] in texts[j].lower():  # cheap prefilter; the reference decides
                has[i, j] = er.para_has_answer(answers[i], texts[j])
    np.savez_compressed(os.path.join(HERE, "eval_fixture.npz"), xb=xb, xq=xq, I=dump["I"].astype(np.int32),
                        D=dump["D"].astype(np.float32), has_answer=np.packbits(has, axis=1), topk=np.int32(K),
                        recall_lines=np.array(lines))
    print("\n".join(lines))
    print("eval_fixture.npz written; seed", seed, "bytes", os.path.getsize(os.path.join(HERE, "eval_fixture.npz")))

    make_kmeans_fixture()
    make_trec_fixture()

    # ---- hand-checkable known answers ---------------------------------------------------------------------------
    cases = []
    # case 1: corpus row j = (j+1) * e_{j % 128}; query = e_0 + 2 e_1.  <q, x_j> = (j+1) if j%128==0, 2(j+1) if j%128==1.
    n = 300
    xb1 = np.zeros((n, D), np.float32)
    for j in range(n):
        xb1[j, j % D] = j + 1
    q1 = np.zeros((1, D), np.float32)
    q1[0, 0], q1[0, 1] = 1, 2
    #   non-zero scores: j=0:1, 128:129, 256:257 ; j=1:4, 129:260, 257:516 ; all others 0 (ties -> lowest ids)
    #   (k = 6 stops before the zero-score ties: which tied ids FAISS keeps there depends on its heap's internal order)
    cases.append(dict(name="axis_aligned_ip", metric=0, k=6, xb_rule="row j = (j+1)*e_(j%128), n=300", n=n,
                      xq=q1.tolist(), I=[[257, 129, 256, 128, 1, 0]], D=[[516, 260, 257, 129, 4, 1]]))
    # case 2: same corpus, L2: |q|^2 = 5, |x_j|^2 = (j+1)^2 ; d = 5 + (j+1)^2 - 2<q,x_j>
    #   j=0: 5+1-2=4 ; j=1: 5+4-8=1 ; j=2: 5+9=14 ; j=3: 5+16=21 ; j=4: 30 ...
    cases.append(dict(name="axis_aligned_l2", metric=1, k=5, xb_rule="row j = (j+1)*e_(j%128), n=300", n=n,
                      xq=q1.tolist(), I=[[1, 0, 2, 3, 4]], D=[[1, 4, 14, 21, 30]]))
    # case 3: k > ntotal padding (IndexFlat::search fills the heap with -FLT_MAX / -1)
    cases.append(dict(name="k_gt_ntotal", metric=0, k=6, xb_rule="row j = (j+1)*e_(j%128), n=3", n=3,
                      xq=q1.tolist(), I=[[1, 0, 2, -1, -1, -1]],
                      D=[[4, 1, 0, -3.4028234663852886e38, -3.4028234663852886e38, -3.4028234663852886e38]]))
    # case 4: all-equal scores: strict '>' replacement while scanning ids upwards keeps the first k ids
    cases.append(dict(name="all_ties_keep_lowest_ids", metric=0, k=4, xb_rule="every row = e_5, n=1000", n=1000,
                      xq=np.eye(1, D, 5, dtype=np.float32).tolist(), I_set=[[0, 1, 2, 3]], D=[[1, 1, 1, 1]]))
    json.dump(cases, open(os.path.join(HERE, "known_answers.json"), "w"), indent=1)
    print("known_answers.json written")


def make_kmeans_fixture():
    """kmeans_fixture.npz — the reference's retrieval/group_paras.py run UNMODIFIED (subprocess, cwd laid out as the script
    expects: encodings/train_para_embed.npy, ../data/retrieve_train.txt, output ../data/data_splits/) with ``import faiss``
    bound to the FAISS restatement (IndexFlatL2 + Clustering + vector_float_to_array).  Stored: the fp16 points, the CLI
    arguments, the final assignment I the script obtained and the split files it wrote (line numbers per split)."""
    n, k, niter = 1200, 10, 6
    rng = np.random.default_rng(77)
    centers = (rng.standard_normal((k, D)) * 4.0).astype(np.float32)
    lab = rng.integers(0, k, size=n)
    from oracle import oracle
    # FAISS seeds the centroids with the first k entries of rand_perm(n, seed 1234 + 1): give each of them its own blob, so
    # every blob owns exactly one centroid throughout and no assignment depends on fp32 rounding
    lab[oracle.rand_perm(n, 1235)[:k]] = np.arange(k)
    x = (centers[lab] + 0.05 * rng.standard_normal((n, D))).astype(np.float16)   # well separated: assignments are unambiguous
    tmp = tempfile.mkdtemp(prefix="proqa_golden_km_")
    os.makedirs(os.path.join(tmp, "retrieval", "encodings"))
    os.makedirs(os.path.join(tmp, "data"))
    np.save(os.path.join(tmp, "retrieval", "encodings", "train_para_embed.npy"), x)
    with open(os.path.join(tmp, "data", "retrieve_train.txt"), "w") as f:
        for i in range(n):
            f.write(f"line {i}\n")
    env = dict(os.environ)
    env["PYTHONPATH"] = os.path.join(HERE, "_oracle_faiss") + os.pathsep + env.get("PYTHONPATH", "")
    env["PROQA_GOLDEN_DUMP"] = os.path.join(tmp, "dump.npz")
    subprocess.run([sys.executable, os.path.join(REF, "group_paras.py"), "--ncentroids", str(k), "--niter", str(niter),
                    "--max_points_per_centroid", "1000"], cwd=os.path.join(tmp, "retrieval"), env=env, capture_output=True, text=True, check=True)
    dump = np.load(os.path.join(tmp, "dump.npz"))
    splits = []
    for c in range(k):
        lines = open(os.path.join(tmp, "data", "data_splits", f"split_{c}.txt")).read().splitlines()
        splits.append(np.array([int(ln.split()[1]) for ln in lines], dtype=np.int32))
    assert sum(len(sp) for sp in splits) == n
    np.savez_compressed(os.path.join(HERE, "kmeans_fixture.npz"), x=x, k=np.int32(k), niter=np.int32(niter), max_points_per_centroid=np.int32(1000),
                        I=dump["I"].astype(np.int32), D=dump["D"].astype(np.float32), split_sizes=np.array([len(sp) for sp in splits], np.int32),
                        split_lines=np.concatenate(splits))
    print("kmeans_fixture.npz written, bytes", os.path.getsize(os.path.join(HERE, "kmeans_fixture.npz")), "split sizes", [len(sp) for sp in splits])


TREC_N, TREC_NQ, TREC_K, TREC_SEED = 12000, 4, 10000, 987


def trec_inputs():
    """Deterministic inputs of the trec fixture (fp16 values, as get_embed.py --fp16 writes them)."""
    rng = np.random.default_rng(TREC_SEED)
    xb = rng.standard_normal((TREC_N, D)).astype(np.float16)
    xq = rng.standard_normal((TREC_NQ, D)).astype(np.float16)
    return xb, xq


def make_trec_fixture():
    import hashlib
    xb, xq = trec_inputs()
    S = xq.astype(np.float64) @ xb.astype(np.float64).T
    order = np.argsort(-S, axis=1)
    # labels: queries 0, 1, 3 have a relevant passage well inside the top 10000 (true ranks 3, 700, 9000), query 2 only
    # far outside it (rank 11500 of 12000) — so the printed recall is 0.75 and no label sits near the k-th place
    labels = [[int(order[0, 3]), int(order[0, 11000])], [int(order[1, 700])], [int(order[2, 11500])], [int(order[3, 9000])]]
    tmp = tempfile.mkdtemp(prefix="proqa_golden_trec_")
    np.save(os.path.join(tmp, "paras.npy"), xb)
    np.save(os.path.join(tmp, "queries.npy"), xq)
    with open(os.path.join(tmp, "queries.txt"), "w") as f:
        for i in range(TREC_NQ):
            f.write(json.dumps({"question": f"q {i}", "labels": labels[i], "qid": 100 + i}) + "\n")
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(HERE, "_oracle_faiss"), REF, env.get("PYTHONPATH", "")])
    code = ("import trec_process as t; t.retrieve_topk(index_path='paras.npy', query_embeds='queries.npy', "
            "query_input='queries.txt', output='out.txt')")
    out = subprocess.run([sys.executable, "-c", code], cwd=tmp, env=env, capture_output=True, text=True, check=True)
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("Avg recall")]
    assert len(line) == 1, out.stdout + out.stderr
    rows = [json.loads(ln) for ln in open(os.path.join(tmp, "out.txt"))]
    I = np.array([r["para_embed_idx"] for r in rows], dtype=np.int32)
    assert I.shape == (TREC_NQ, TREC_K)
    n_hit = np.array([int(np.sum(r["para_labels"])) for r in rows], np.int32)
    np.savez_compressed(os.path.join(HERE, "trec_fixture.npz"), seed=np.int32(TREC_SEED), n=np.int32(TREC_N), nq=np.int32(TREC_NQ),
                        k=np.int32(TREC_K), xb_sha1=np.array(hashlib.sha1(xb.tobytes()).hexdigest()),
                        xq_sha1=np.array(hashlib.sha1(xq.tobytes()).hexdigest()), I=I, label_hits=n_hit,
                        labels=np.array([json.dumps(lb) for lb in labels]), recall_line=np.array(line[0]))
    print(line[0], "| label hits", n_hit.tolist(), "| trec_fixture.npz bytes", os.path.getsize(os.path.join(HERE, "trec_fixture.npz")))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "kmeans":
        make_kmeans_fixture()
    elif len(sys.argv) > 1 and sys.argv[1] == "trec":
        make_trec_fixture()
    else:
        main()
