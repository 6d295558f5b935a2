#!/usr/bin/env python
"""TEST INFRASTRUCTURE.  Full-size golden vectors for the benchmarked workloads.

Generates the synthetic collection of a bench.py workload with bench.py's own generator and seed, runs the reference build
(oracle/_ref/mumemto_exec = the reference's sources compiled unmodified, oracle/Makefile) on it with the given flags, and records the
md5 / size / line count of every output in tests/golden/md5_table.json under the key  full:<workload>:<flagset>.  bench.py compares the
md5 of its own output with that entry off the clock.  Runs in the build container only (needs /root/reference to have built oracle/_ref;
the GPU box never runs this).

  python scripts/make_fullsize_golden.py --workload c2 --flagset strict [--keep] [--dir /tmp/fullsize]
"""
import argparse
import hashlib
import json
import os
import resource
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

FLAGSETS = {"strict": [], "partial": ["-k", "-1"], "mem": ["-f", "3", "-F", "100"], "anchor": ["-M", "-n"]}
OUT_EXT = [".mums", ".mems", ".athresh"]


def md5_of(path):
    h = hashlib.md5()
    lines = 0
    with open(path, "rb") as f:
        while True:
            b = f.read(1 << 24)
            if not b:
                break
            h.update(b)
            lines += b.count(b"\n")
    return h.hexdigest(), lines


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", required=True)
    ap.add_argument("--flagset", default="strict")
    ap.add_argument("--dir", default="/tmp/fullsize")
    ap.add_argument("--keep", action="store_true", help="keep the output files (for a byte-wise diff with a GPU run)")
    a = ap.parse_args()
    spec = bench.WORKLOADS[a.workload]
    d = os.path.join(a.dir, a.workload)
    os.makedirs(d, exist_ok=True)
    paths = [os.path.join(d, f"h{i}.fa") for i in range(spec["N"])]
    if not all(os.path.exists(p) for p in paths):
        t0 = time.time()
        for p, im in zip(paths, bench.make_collection(spec, seed=1)):
            im.tofile(p)
        print(f"generated {a.workload} in {time.time() - t0:.0f} s", flush=True)
    exe = os.path.join(ROOT, "oracle", "_ref", "mumemto_exec")
    prefix = os.path.join(d, f"ref_{a.flagset}")
    t0 = time.time()
    subprocess.run([exe, "-o", prefix, *FLAGSETS[a.flagset], *paths], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    secs = time.time() - t0
    peak_gb = resource.getrusage(resource.RUSAGE_CHILDREN).ru_maxrss / 1e6
    entry = {"generator": f"bench.make_collection(WORKLOADS['{a.workload}'], seed=1)", "flags": FLAGSETS[a.flagset], "reference_seconds": round(secs, 1),
             "reference_peak_rss_gb": round(peak_gb, 2), "input_bases": spec["N"] * spec["L"], "md5": {}, "bytes": {}, "lines": {}}
    for ext in OUT_EXT:
        p = prefix + ext
        if os.path.exists(p):
            entry["md5"][ext], entry["lines"][ext] = md5_of(p)
            entry["bytes"][ext] = os.path.getsize(p)
            if ext == ".athresh":
                del entry["lines"][ext]
            if not a.keep:
                os.remove(p)
    table_path = os.path.join(ROOT, "tests", "golden", "md5_table.json")
    with open(table_path) as f:
        table = json.load(f)
    table[f"full:{a.workload}:{a.flagset}"] = entry
    with open(table_path, "w") as f:
        json.dump(table, f, indent=1, sort_keys=True)
        f.write("\n")
    print(json.dumps(entry), flush=True)


if __name__ == "__main__":
    main()
