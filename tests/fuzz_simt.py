"""Differential fuzzing of the engine (SIMT unit-test build) against the CPU oracle: random small collections, random flags, both routes,
random tile sizes.  Not collected by pytest (no test_ prefix); run it by hand:  python tests/fuzz_simt.py [iterations] [seed]"""
import ctypes
import os
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import util  # noqa: E402
from mumemto_b200 import _binding, api  # noqa: E402

ORACLE = os.path.join(ROOT, "oracle", "_build", "mumemto_oracle")
REF = os.path.join(ROOT, "oracle", "_ref", "mumemto_exec")


def random_collection(rng, d):
    n = int(rng.integers(2, 7))
    base_len = int(rng.integers(40, 2500))
    if rng.random() < 0.25:           # many similar documents: multi-member dictionary groups, multi-word document masks
        n = int(rng.integers(7, 75)); base_len = int(rng.integers(150, 900))
    base = util.random_dna(rng, base_len)
    if rng.random() < 0.4:            # tandem repeats / low complexity
        unit = util.random_dna(rng, int(rng.integers(1, 12)))
        at = int(rng.integers(0, base_len))
        base = base[:at] + unit * int(rng.integers(2, 60)) + base[at:]
    paths = []
    for i in range(n):
        s = util.mutate(rng, base, float(rng.choice([0.0, 0.002, 0.01, 0.05])))
        if rng.random() < 0.3:
            cut = int(rng.integers(0, len(s)))
            s = s[cut:] + s[:cut]
        if rng.random() < 0.2:
            at = int(rng.integers(0, len(s)))
            s = s[:at] + "N" * int(rng.integers(1, 30)) + s[at:]
        if rng.random() < 0.15:
            s = s.lower()
        if rng.random() < 0.2:        # reverse complement of the whole document
            s = s[::-1].translate(str.maketrans("ACGTacgt", "TGCAtgca"))
        recs = [(f"r{i}", s)]
        if rng.random() < 0.25 and len(s) > 20:
            cut = int(rng.integers(1, len(s) - 1))
            recs = [(f"r{i}a", s[:cut]), (f"r{i}b", s[cut:])]
        paths.append(util.write_fasta(os.path.join(d, f"d{i}.fa"), recs, width=int(rng.choice([60, 80, 7, 1000]))))
    return paths


def random_flags(rng, n):
    mode = rng.integers(0, 6) if rng.random() > float(os.environ.get("FUZZ_MERGE_SHARE", "0")) else 3
    flags = []
    if mode == 0:
        pass
    elif mode == 1:
        flags += ["-k", str(int(rng.integers(-n - 1, n + 2)))]
    elif mode == 2:
        flags += ["-f", str(int(rng.integers(2, 5)))]
        if rng.random() < 0.5:
            flags += ["-F", str(int(rng.integers(-n - 1, 4 * n)))]
        if rng.random() < 0.5:
            flags += ["-k", str(int(rng.integers(2, n + 1)))]
    elif mode == 3:
        flags += ["-M", "-n"]
    elif mode == 4:
        flags += ["-f", "0"]
        if rng.random() < 0.6:
            flags += ["-F", str(int(rng.integers(2, 40)))]
        flags += ["-k", str(int(rng.integers(2, n + 1)))]
    else:
        flags += ["-b"] if rng.random() < 0.5 else ["-A"]
    if rng.random() < 0.3:
        flags += ["-l", str(int(rng.integers(3, 40)))]
    if rng.random() < 0.25:
        flags += ["-r"]
    return flags


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    skip = int(sys.argv[3]) if len(sys.argv) > 3 else 0           # replay: draw the first `skip` cases without running them
    keep = os.environ.get("FUZZ_KEEP")                            # directory that receives the inputs of every case before it runs
    lib = _binding.declare(ctypes.CDLL(os.path.join(ROOT, "tests", "simt", "_build", "libmumemto_b200_simt.so")))
    rng = np.random.default_rng(seed)
    bad = 0
    for it in range(iters):
        d = tempfile.mkdtemp(prefix="fuzz")
        try:
            paths = random_collection(rng, d)
            flags = random_flags(rng, len(paths))
            o, g = os.path.join(d, "o"), os.path.join(d, "g")
            route = str(rng.choice(["direct", "pfp", "auto"]))
            tile = int(rng.choice([0, 64, 300, 1500, 100000]))
            if it < skip:
                continue
            if keep:
                shutil.rmtree(keep, ignore_errors=True); shutil.copytree(d, keep)
                with open(os.path.join(keep, "case.txt"), "w") as f:
                    f.write(f"{it} {flags} {route} {tile}\n")
            os.environ["MMB_DOC_STREAM"] = "1" if rng.random() < 0.5 else "0"
            # dictionary order of the prefix-free-parse route: chunk engine with segment pruning (default), the same without pruning,
            # the doubling engine (MMB_NO_PRUNE), and the fallback from the chunk engine after too few rounds
            variant = rng.random()
            no_prune = bool(variant < 0.15)
            for k in ("MMB_NO_PRUNE", "MMB_CHUNK_ROUNDS", "MMB_THRESH_SLOW"):
                os.environ.pop(k, None)
            if rng.random() < 0.3:
                os.environ["MMB_THRESH_SLOW"] = "1"         # -M thresholds of pruned segments from the occurrence lists instead of the phrase masks
            if 0.15 <= variant < 0.25:
                os.environ["MMB_NO_PRUNE"] = "1"
            elif 0.25 <= variant < 0.35:
                os.environ["MMB_CHUNK_ROUNDS"] = str(int(rng.integers(1, 4)))
            lib.simt_set_block_order(int(rng.integers(0, 3)))           # block scheduling order of the SIMT harness: forward, reverse, shuffled
            # host-image loads in small upload groups, with the trigger scan running ahead of the parse on the finished text prefix
            for k in ("MMB_UPLOAD_GROUP_BYTES", "MMB_TRIGGER_AHEAD_MIN"):
                os.environ.pop(k, None)
            if rng.random() < 0.3:
                os.environ["MMB_UPLOAD_GROUP_BYTES"] = str(int(rng.integers(200, 6000))); os.environ["MMB_TRIGGER_AHEAD_MIN"] = "0"
            if os.path.exists(REF) and rng.random() < 0.15 and "-A" not in flags and not ("-f" in flags and flags[flags.index("-f") + 1] == "0" and "-F" not in flags):
                # array input: the reference's own -A dump, then its own -a run against ours
                import subprocess
                full = os.path.join(d, "full")
                rc = [f for f in flags if f == "-r"]
                if subprocess.run([REF, "-o", full, "-A", *rc, *paths], capture_output=True).returncode == 0:
                    ok_o = subprocess.run([REF, "-o", o, "-a", full, *flags], capture_output=True).returncode == 0
                    try:
                        api.find_arrays(full, g, tile_ranks=tile, _lib=lib, **util.flags_to_kwargs(flags)); ok_g = True
                    except Exception as ex:
                        ok_g = False; err = str(ex)
                    if ok_o != ok_g or (ok_o and util.outputs(g) != util.outputs(o)):
                        bad += 1; print("MISMATCH arrays", it, flags, tile, ok_o, ok_g); shutil.copytree(d, f"/tmp/fuzz_fail_{seed}_{it}")
                continue
            if len(paths) >= 3 and rng.random() < 0.12:
                # partition -> anchor merge: random partitions sharing document 0, ours against the oracle's fold
                k = int(rng.integers(2, min(4, len(paths) - 1) + 1))
                rest = list(rng.permutation(np.arange(1, len(paths))))
                groups = [[0] + [int(x) for x in rest[j::k]] for j in range(k)]
                extra = ["-r"] if "-r" in flags else []
                no, ng = [], []
                try:
                    for j, grp in enumerate(groups):
                        sel = [paths[x] for x in grp]
                        util.run_cli(ORACLE, ["-o", os.path.join(d, f"po{j}"), "-M", "-n", *extra, *sel])
                        api.find_files(sel, os.path.join(d, f"pg{j}"), merge=True, anchor_merge=True, revcomp=not extra, route=route, tile_ranks=tile, no_prune=no_prune, _lib=lib)
                        no.append(os.path.join(d, f"po{j}.mums")); ng.append(os.path.join(d, f"pg{j}.mums"))
                    util.run_cli(ORACLE, ["merge", *no, "-o", os.path.join(d, "om")])
                    api.merge_files(ng, os.path.join(d, "gm"), _lib=lib)
                    if util.outputs(os.path.join(d, "gm")) != util.outputs(os.path.join(d, "om")):
                        bad += 1; print("MISMATCH merge", it, groups, route, tile); shutil.copytree(d, f"/tmp/fuzz_fail_{seed}_{it}")
                except Exception as ex:
                    if "too short" not in str(ex):
                        bad += 1; print("ERROR merge", it, groups, route, tile, str(ex)[:300]); shutil.copytree(d, f"/tmp/fuzz_fail_{seed}_{it}")
                continue
            ok_o = util.run_cli(ORACLE, ["-o", o, *flags, *paths], check=False).returncode == 0
            unbounded = "-f" in flags and flags[flags.index("-f") + 1] == "0" and "-F" not in flags
            if unbounded and route == "pfp":
                # one view of the PRUNED stream: only when pruning is on and possible (-l above 10, chunk engine not forced to fall back)
                minlen = int(flags[flags.index("-l") + 1]) if "-l" in flags else 20
                if no_prune or minlen <= 10 or os.environ.get("MMB_NO_PRUNE") or os.environ.get("MMB_CHUNK_ROUNDS"):
                    route = "direct"
            kw = util.flags_to_kwargs(flags)
            try:
                api.find_files(paths, g, route=route, tile_ranks=tile, no_prune=no_prune, _lib=lib, **kw); ok_g = True
            except Exception as ex:
                ok_g = False; err = str(ex)
            if ok_o != ok_g:
                too_short = (not ok_g) and "too short" in err
                if not too_short:
                    bad += 1; print("MISMATCH status", it, flags, route, tile, ok_o, ok_g, "" if ok_g else err[:200]); shutil.copytree(d, f"/tmp/fuzz_fail_{seed}_{it}")
            elif ok_o and util.outputs(g) != util.outputs(o):
                bad += 1; print("MISMATCH output", it, flags, route, tile, no_prune, os.environ.get("MMB_NO_PRUNE"), os.environ.get("MMB_CHUNK_ROUNDS")); shutil.copytree(d, f"/tmp/fuzz_fail_{seed}_{it}")
            elif ok_o and rng.random() < 0.25 and not any(f in flags for f in ("-M", "-A", "-b")) and not no_prune:
                # one collection in K slices of the suffix order: the slices' outputs concatenate to the whole
                K = int(rng.integers(2, 6)); ext = ".mems" if "-f" in flags and flags[flags.index("-f") + 1] != "1" else ".mums"
                try:
                    got = b""
                    for i in range(K):
                        api.find_files(paths, g + f"_s{i}", tile_ranks=tile, shard=(i, K), _lib=lib, **kw)
                        got += util.outputs(g + f"_s{i}").get(ext, b"")
                    if got != util.outputs(o).get(ext, b""):
                        bad += 1; print("MISMATCH sharded", it, flags, K, tile); shutil.copytree(d, f"/tmp/fuzz_fail_{seed}_{it}")
                except Exception as ex:
                    if "sharded call needs" not in str(ex) and "too short" not in str(ex):
                        bad += 1; print("ERROR sharded", it, flags, K, str(ex)[:200]); shutil.copytree(d, f"/tmp/fuzz_fail_{seed}_{it}")
        finally:
            shutil.rmtree(d, ignore_errors=True)
        if (it + 1) % 50 == 0:
            print(f"{it + 1} iterations, {bad} mismatches", flush=True)
    print("done:", iters, "iterations,", bad, "mismatches")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
