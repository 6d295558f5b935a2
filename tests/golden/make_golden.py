#!/usr/bin/env python
"""Regenerate tests/golden/: run in the development container (needs oracle/_ref, i.e. /root/reference).

Inputs are produced from seeds by tests/util.appendix_a; outputs come from oracle/_ref/mumemto_exec and
oracle/_ref/anchor_merge = the reference's own sources compiled unmodified against the stand-ins in oracle/ref_shims/.
Every md5 goes into md5_table.json; the small t2 text outputs are committed as files too.
"""
import json
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import util  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")


def run(exe, args, cwd):
    subprocess.run([os.path.join(REF, exe), *args], cwd=cwd, check=True, capture_output=True)


def main():
    table = {}
    tmp = tempfile.mkdtemp()
    sets = {"t2": dict(N=5, L=20000, seed=42, identical_ends=True), "t7": dict(N=5, L=1000000, seed=1)}
    flagsets = {"strict": [], "partial": ["-k", "-1"], "mem": ["-f", "3", "-F", "100"], "bumbl": ["-b"], "partial_bumbl": ["-k", "-1", "-b"],
                "anchor": ["-M", "-n"], "string_merge": ["-M"], "arrays": ["-A"], "norc": ["-r"], "short": ["-l", "12"],
                "mem2": ["-k", "3", "-f", "2"], "memF": ["-f", "0", "-F", "7", "-k", "2"]}
    for name, kw in sets.items():
        d = os.path.join(tmp, name)
        paths = util.appendix_a(d, **kw)
        table[name] = {"inputs": {os.path.basename(p): util.md5_file(p) for p in paths}, "outputs": {}}
        for fs, flags in flagsets.items():
            if name == "t7" and fs not in ("strict", "partial", "mem", "anchor"):
                continue
            run("mumemto_exec", ["-o", fs, *flags, *[os.path.basename(p) for p in paths]], d)
            outs = util.outputs(os.path.join(d, fs))
            table[name]["outputs"][fs] = {"flags": flags, "md5": {ext: util.md5(b) for ext, b in outs.items()}}
            if name == "t2":
                for ext, b in outs.items():
                    if ext in (".mums", ".mems") and len(b) < 200000:
                        os.makedirs(os.path.join(HERE, "t2"), exist_ok=True)
                        with open(os.path.join(HERE, "t2", fs + ext), "wb") as f:
                            f.write(b)
        if name == "t2":  # partition merge (t2-merge, t2-merge3)
            b = [os.path.basename(p) for p in paths]
            for tag, parts in {"merge2": [[0, 1, 2], [0, 3, 4]], "merge3": [[0, 1], [0, 2, 3], [0, 4]]}.items():
                names = []
                for i, idx in enumerate(parts):
                    run("mumemto_exec", ["-o", f"{tag}_p{i}", "-M", "-n", *[b[j] for j in idx]], d)
                    names.append(f"{tag}_p{i}.mums")
                run("anchor_merge", [*names, "-o", tag], d)
                outs = util.outputs(os.path.join(d, tag))
                table[name]["outputs"][tag] = {"partitions": parts, "md5": {ext: util.md5(x) for ext, x in outs.items()}}
                with open(os.path.join(HERE, "t2", tag + ".mums"), "wb") as f:
                    f.write(outs[".mums"])
    # toy from the reference's library README (mumemto_library/README.md:53,58)
    d = os.path.join(tmp, "toy"); os.makedirs(d)
    util.write_fasta(os.path.join(d, "a.fa"), [("a", "ACGTACGT")]); util.write_fasta(os.path.join(d, "b.fa"), [("b", "TACGTAAA")])
    run("mumemto_exec", ["-o", "toy", "-l", "3", "-f", "2", "a.fa", "b.fa"], d)
    with open(os.path.join(d, "toy.mems"), "rb") as f:
        toy = f.read()
    table["toy"] = {"mems": toy.decode(), "md5": util.md5(toy)}
    with open(os.path.join(HERE, "md5_table.json"), "w") as f:
        json.dump(table, f, indent=1, sort_keys=True)
    shutil.rmtree(tmp)
    print(json.dumps({k: list(v.get("outputs", {}).keys()) for k, v in table.items()}))


if __name__ == "__main__":
    main()
