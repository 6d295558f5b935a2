#!/usr/bin/env python
"""Regenerate tests/golden/string_merge/: run in the development container (needs /root/reference and oracle/_ref).

For every case: seeded FASTA inputs -> the reference's `mumemto_exec -M` per partition (oracle/_ref, its own sources) -> the
reference's own `mumemto/merge_mums.py` (run as a script with the reference's `extract_mums` / `mumemto_exec` on PATH).  The inputs of
the merge step proper (partition `.mums/.thresh/.thresh_rev`, the MUMs-of-MUMs `.mums/.lengths`) and its outputs are committed, so the
CPU test checks `mumemto_b200.merge_mums.merge_tables` without the reference and the GPU test checks the whole workflow on our
binaries.  `split_vectors.json` holds input/output pairs of the reference's `remove_start_dollar` on random cases (imported function).
"""
import importlib.util
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import util  # noqa: E402

REF_BIN = os.path.join(ROOT, "oracle", "_ref")
REF_PY = "/root/reference/mumemto/merge_mums.py"
OUT = os.path.join(HERE, "string_merge")

COMP = str.maketrans("ACGT", "TGCA")


def case_inputs(d, name):
    """FASTA files + partition index lists of a case; same generator as the GPU test (tests/test_merge_mums.py imports this)."""
    if name == "two":          # 6 haplotypes, 2 partitions
        return util.appendix_a(d, N=6, L=8000, seed=7), [[0, 1, 2], [3, 4, 5]]
    if name == "three":        # 7 haplotypes, 3 partitions of unequal size
        return util.appendix_a(d, N=7, L=6000, seed=11, snp=0.006), [[0, 1], [2, 3, 4], [5, 6]]
    if name == "reverse":      # the second partition holds reverse-complemented haplotypes: every MUM-of-MUMs is a '-' match there
        paths = util.appendix_a(d, N=4, L=6000, seed=5, snp=0.008)
        for p in paths[2:]:
            lines = open(p).read().split("\n")
            seq = "".join(lines[1:])[::-1].translate(COMP)
            util.write_fasta(p, [(lines[0][1:], seq)], width=80)
        return paths, [[0, 1], [2, 3]]
    raise KeyError(name)


CASES = ["two", "three", "reverse"]


def main():
    shutil.rmtree(OUT, ignore_errors=True)
    os.makedirs(OUT)
    env = dict(os.environ, PATH=REF_BIN + ":" + os.environ["PATH"], PYTHONPATH="/root/reference")
    table = {}
    for name in CASES:
        d = tempfile.mkdtemp()
        paths, parts = case_inputs(d, name)
        b = [os.path.basename(p) for p in paths]
        prefixes = []
        for i, idx in enumerate(parts):
            subprocess.run([os.path.join(REF_BIN, "mumemto_exec"), "-o", f"p{i}", "-M", *[b[j] for j in idx]], cwd=d, check=True, capture_output=True)
            prefixes.append(f"p{i}")
        # the intermediate files are removed by the script: run the two steps by hand first to keep them, then the script itself
        for p in prefixes:
            subprocess.run([os.path.join(REF_BIN, "extract_mums"), "-m", p + ".mums"], cwd=d, check=True)
        subprocess.run([os.path.join(REF_BIN, "mumemto_exec"), *[p + "_mums.fa" for p in prefixes], "-o", "mom"], cwd=d, check=True, capture_output=True)
        r = subprocess.run([sys.executable, REF_PY, *[p + ".mums" for p in prefixes], "-o", "merged"], cwd=d, env=env, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(r.stderr[-3000:])
        r = subprocess.run([sys.executable, REF_PY, "-m", "mom.mums", *[p + ".mums" for p in prefixes], "-o", "merged_m.bumbl"], cwd=d, env=env, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(r.stderr[-3000:])
        for p in prefixes:                                    # the script removed them again
            subprocess.run([os.path.join(REF_BIN, "extract_mums"), "-m", p + ".mums"], cwd=d, check=True)
        o = os.path.join(OUT, name)
        os.makedirs(o)
        keep = [p + e for p in prefixes for e in (".mums", ".thresh", ".thresh_rev", "_mums.fa")] + \
               ["mom.mums", "merged.mums", "merged.thresh", "merged.thresh_rev", "merged_m.bumbl"]
        for f in keep:
            shutil.copy(os.path.join(d, f), os.path.join(o, f))
        with open(os.path.join(d, "mom.lengths")) as f, open(os.path.join(o, "mom.lengths"), "w") as g:     # paths -> base names
            for line in f:
                t = line.split()
                g.write(" ".join([os.path.basename(t[0])] + t[1:]) + "\n")
        table[name] = {"partitions": parts, "md5": {f: util.md5_file(os.path.join(o, f)) for f in sorted(os.listdir(o))},
                       "merged_rows": sum(1 for _ in open(os.path.join(o, "merged.mums")))}
        shutil.rmtree(d)

    spec = importlib.util.spec_from_file_location("ref_merge_mums", REF_PY)
    sys.path.insert(0, "/root/reference/mumemto")      # `from utils import ...` resolves without the compiled package
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    rng = np.random.default_rng(3)
    vectors = []
    for _ in range(60):
        K = int(rng.integers(2, 5))
        recs = rng.integers(21, 60, int(rng.integers(3, 9)))
        ends = np.cumsum(recs)
        bv = np.zeros(int(ends[-1]) + 1, dtype=bool)
        bv[ends - 1] = True
        l = int(rng.integers(20, 120))
        s0 = int(rng.integers(0, max(1, int(ends[-1]) - l)))
        starts = [s0] + [int(x) for x in rng.integers(0, 1000, K - 1)]
        strands = [bool(x) for x in rng.integers(0, 2, K)]
        out = ref.remove_start_dollar((l, starts, strands), bv)
        vectors.append({"records": [int(x) for x in recs], "l": l, "starts": starts, "strands": strands,
                        "pieces": [[int(a), [int(x) for x in b], [bool(x) for x in c]] for a, b, c in out]})
    with open(os.path.join(OUT, "split_vectors.json"), "w") as f:
        json.dump(vectors, f)
    with open(os.path.join(OUT, "table.json"), "w") as f:
        json.dump(table, f, indent=1, sort_keys=True)
    print(json.dumps({k: v["merged_rows"] for k, v in table.items()}))


if __name__ == "__main__":
    main()
