"""Shared helpers for the tests: seeded synthetic inputs (SURVEY.md Appendix A generator), runners, comparisons."""
import hashlib
import json
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def md5(b):
    return hashlib.md5(b).hexdigest()


def md5_file(p):
    with open(p, "rb") as f:
        return md5(f.read())


def appendix_a(outdir, N, L, seed, identical_ends=False, snp=0.01, width=80):
    """The generator of SURVEY.md Appendix A (t2: N=5, L=20000, seed=42, identical_ends; t7 = C1: N=5, L=1e6, seed=1)."""
    os.makedirs(outdir, exist_ok=True)
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 4, L)
    paths = []
    for i in range(N):
        s = base.copy()
        m = rng.random(L) < snp
        if identical_ends:
            m[:40] = False
            m[-40:] = False
        s[m] = (s[m] + rng.integers(1, 4, m.sum())) % 4
        seq = ACGT[s].tobytes().decode()
        p = os.path.join(outdir, f"h{i}.fa")
        with open(p, "w") as f:
            f.write(f">h{i}\n" + "".join(seq[k:k + width] + "\n" for k in range(0, L, width)))
        paths.append(p)
    return paths


def write_fasta(path, records, width=60, newline="\n"):
    with open(path, "w", newline="") as f:
        for name, seq in records:
            f.write(f">{name}{newline}")
            for k in range(0, len(seq), width):
                f.write(seq[k:k + width] + newline)
    return path


def random_dna(rng, n, alphabet="ACGT"):
    a = np.frombuffer(alphabet.encode(), dtype=np.uint8)
    return a[rng.integers(0, len(a), n)].tobytes().decode()


def mutate(rng, seq, rate):
    s = np.frombuffer(seq.encode(), dtype=np.uint8).copy()
    m = rng.random(len(s)) < rate
    s[m] = ACGT[rng.integers(0, 4, m.sum())]
    return s.tobytes().decode()


def run_cli(binary, args, cwd=None, check=True):
    r = subprocess.run([binary, *map(str, args)], cwd=cwd, capture_output=True)
    if check and r.returncode != 0:
        raise RuntimeError(f"{binary} {' '.join(map(str, args))} failed ({r.returncode}):\n{r.stderr.decode(errors='replace')[-2000:]}")
    return r


FLAG_SETS = {
    "strict": [],
    "partial": ["-k", "-1"],
    "mem": ["-f", "3", "-F", "100"],
    "norc": ["-r"],
    "short": ["-l", "12"],
    "mem2": ["-k", "3", "-f", "2"],
    "memF": ["-f", "0", "-F", "7", "-k", "2"],
}


def flags_to_kwargs(flags):
    """CLI flags -> keyword arguments of mumemto_b200.api.find_files."""
    kw = {}
    it = iter(flags)
    for a in it:
        if a == "-k": kw["num_distinct_docs"] = int(next(it))
        elif a == "-f": kw["rare_freq"] = int(next(it))
        elif a == "-F": kw["max_mem_freq"] = int(next(it))
        elif a == "-l": kw["min_match_len"] = int(next(it))
        elif a == "-r": kw["revcomp"] = False
        elif a == "-b": kw["binary"] = True
        elif a == "-M": kw["merge"] = True
        elif a == "-n": kw["anchor_merge"] = True; kw["merge"] = True
        elif a == "-A": kw["arrays_out"] = True
        else: raise ValueError(a)
    return kw


def outputs(prefix):
    """{extension: bytes} of every output file with this prefix (lengths excluded: it embeds absolute paths)."""
    d, base = os.path.dirname(prefix), os.path.basename(prefix)
    out = {}
    for f in sorted(os.listdir(d)):
        if f.startswith(base + ".") and not f.endswith(".lengths"):
            with open(os.path.join(d, f), "rb") as fh:
                out[f[len(base):]] = fh.read()
    return out


def load_golden_table():
    with open(os.path.join(GOLDEN, "md5_table.json")) as f:
        return json.load(f)
