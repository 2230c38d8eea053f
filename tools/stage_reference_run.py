#!/usr/bin/env python
"""Stage the UNMODIFIED reference scripts and the eval fixture's scratch tree under baseline/_ref/ (git-ignored, but it travels
to the GPU box with the snapshot), so that tests/test_gpu_reference_script.py can run retrieval/eval_retrieval.py itself on a
B200 through the faiss shim.  Run HERE (needs /root/reference); nothing of the reference enters the repository's history.

    python tools/stage_reference_run.py
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
REF = "/root/reference/retrieval"


def main():
    if not os.path.isdir(REF):
        raise SystemExit("no /root/reference here: nothing staged")
    import make_fixtures as mf
    dst = os.path.join(ROOT, "baseline", "_ref")
    run = os.path.join(dst, "eval_run")
    os.makedirs(os.path.join(run, "retrieval"), exist_ok=True)
    for name in ("eval_retrieval.py", "basic_tokenizer.py", "utils.py"):  # (what the script imports)
        shutil.copyfile(os.path.join(REF, name), os.path.join(run, "retrieval", name))
    seed, xb, xq = mf.find_inputs()
    mf.build_eval_tree(run, seed, xb, xq)
    print("staged", run, "seed", seed)


if __name__ == "__main__":
    main()
