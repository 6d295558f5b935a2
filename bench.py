#!/usr/bin/env python
"""bench.py -- input bases/sec of the FASTA -> .mums hot path on B200 (BASELINE.json metric), one process per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c1|c2|...] [--impl ours|reference]

A step = one pass of the hot path over one synthetic collection:
  value : FASTA bytes already resident in HBM -> text build -> suffix array -> LCP/BWT -> interval detection -> formatted .mums bytes
          (left on the device); timed per step with CUDA events on the stream the kernels are launched on, L2 flushed between steps.
  e2e   : the same through the reference-facing call with HOST buffers: pinned FASTA images -> H2D -> hot path -> D2H of the .mums bytes.
N > 1 (torchrun): every rank owns one partition (shared anchor h0 + its own haplotypes, -M -n), no collective on the data path
(scaling "weak") until the exchange step inside every timed step: one NCCL all-gather of the partition tables, then the merge kernels.
--impl reference times the reference's own CPU implementation (oracle/_ref/mumemto_exec = reference sources compiled unmodified,
single-threaded because the reference has no threads) on the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (N docs, L bases per doc, generator, description)
    "c1": dict(N=5, L=1_000_000, kind="independent", snp=0.01, desc="5x1 Mbp synthetic haplotypes (1% SNP), strict multi-MUMs [BASELINE configs[0]]"),
    "c1x4": dict(N=5, L=4_000_000, kind="independent", snp=0.01, desc="5x4 Mbp synthetic haplotypes (1% SNP), strict multi-MUMs"),
    "h64_1m": dict(N=64, L=1_000_000, kind="pangenome", desc="64x1 Mbp human-chr-like haplotypes, strict multi-MUMs"),
    "h64_4m": dict(N=64, L=4_000_000, kind="pangenome", desc="64x4 Mbp human-chr-like haplotypes, strict multi-MUMs"),
    "h64_16m": dict(N=64, L=16_000_000, kind="pangenome", desc="64x16 Mbp human-chr-like haplotypes, strict multi-MUMs"),
    "h64_40m": dict(N=64, L=40_000_000, kind="pangenome", desc="64x40 Mbp human-chr-like haplotypes (text crosses 2^32 symbols), strict multi-MUMs"),
    "h64_40m_c3": dict(N=64, L=40_000_000, kind="pangenome", flagset="partial", desc="64x40 Mbp human-chr-like haplotypes (text crosses 2^32 symbols), partial multi-MUMs -k -1"),
    "h64_40m_c4": dict(N=64, L=40_000_000, kind="pangenome", flagset="mem", desc="64x40 Mbp human-chr-like haplotypes (text crosses 2^32 symbols), multi-MEMs -f 3 -F 100"),
    "c2": dict(N=64, L=100_000_000, kind="pangenome", desc="64x100 Mbp synthetic human-chr-like haplotypes, strict multi-MUMs [BASELINE configs[1]]"),
    "c3": dict(N=64, L=100_000_000, kind="pangenome", flagset="partial", desc="64x100 Mbp synthetic human-chr-like haplotypes, partial multi-MUMs -k -1 [BASELINE configs[2]]"),
    "c4": dict(N=64, L=100_000_000, kind="pangenome", flagset="mem", desc="64x100 Mbp synthetic human-chr-like haplotypes, multi-MEMs -f 3 -F 100 [BASELINE configs[3]]"),
    # one partition of BASELINE configs[4] (256x100 Mbp = 8 partitions of 32 genomes, 1 per GPU): anchor + 31 haplotypes, -M -n
    "c5": dict(N=32, L=100_000_000, kind="pangenome", desc="partition of 32x100 Mbp haplotypes sharing anchor h0, strict multi-MUMs -M -n [BASELINE configs[4]: 1 partition per GPU, NCCL all-gather + merge]"),
}
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def fasta_image(name, codes, width=80):
    """uint8 codes 0..3 -> FASTA file image (header + 80-column lines) as a uint8 array, vectorised."""
    L = len(codes)
    seq = ACGT[codes]
    nfull, rem = divmod(L, width)
    head = f">{name}\n".encode()
    body = np.empty(L + nfull + (1 if rem else 0), dtype=np.uint8)
    if nfull:
        blk = body[: nfull * (width + 1)].reshape(nfull, width + 1)
        blk[:, :width] = seq[: nfull * width].reshape(nfull, width)
        blk[:, width] = 10
    if rem:
        body[nfull * (width + 1): nfull * (width + 1) + rem] = seq[nfull * width:]
        body[-1] = 10
    return np.concatenate([np.frombuffer(head, dtype=np.uint8), body])


def make_collection(spec, seed, anchor_seed=None, first_index=0):
    """Synthetic collection as a list of FASTA images.
    independent: SURVEY.md Appendix A generator (base + independent per-haplotype SNPs; seed 1 reproduces the golden C1 files).
    pangenome  : ancestral sequence (GC ~ 0.41), segregating sites at density 1/300 with derived-allele frequency ~ 1/x on [1/64, 1]
                 shared across haplotypes, plus 0.01 % private SNPs per haplotype."""
    N, L = spec["N"], spec["L"]
    imgs = []
    if spec["kind"] == "independent":
        rng = np.random.default_rng(seed)
        base = rng.integers(0, 4, L)
        for i in range(N):
            s = base.copy()
            m = rng.random(L) < spec["snp"]
            s[m] = (s[m] + rng.integers(1, 4, m.sum())) % 4
            imgs.append(fasta_image(f"h{i}", s.astype(np.uint8)))
        return imgs
    arng = np.random.default_rng(seed if anchor_seed is None else anchor_seed)
    anc = arng.choice(4, size=L, p=[0.295, 0.205, 0.205, 0.295]).astype(np.uint8)
    nsites = L // 300
    sites = np.sort(arng.choice(L, size=nsites, replace=False))
    u = arng.random(nsites)
    freq = (1.0 / 64) * (64.0 ** u)                       # density ~ 1/x on [1/64, 1]
    alt = ((anc[sites] + arng.integers(1, 4, nsites)) % 4).astype(np.uint8)
    rng = np.random.default_rng(seed)
    for i in range(N):
        s = anc.copy()
        if not (anchor_seed is not None and i == 0):      # document 0 of a partition is the shared anchor = the ancestral haplotype pattern
            carry = rng.random(nsites) < freq
            s[sites[carry]] = alt[carry]
            priv = rng.random(L) < 1e-4
            s[priv] = (s[priv] + rng.integers(1, 4, int(priv.sum()))) % 4
        else:
            carry = np.random.default_rng(anchor_seed + 1).random(nsites) < freq
            s[sites[carry]] = alt[carry]
        imgs.append(fasta_image(f"h{first_index + i}", s))
    return imgs


def verify_sample(imgs, text, how_many=400, width=80):
    """Check sampled output lines against the generated sequences (outside the timed region): every occurrence of a multi-MUM must
    spell the same string (reverse-complemented on '-'), and its length must be at least 20.  Returns the number of lines checked."""
    if not text:
        return 0
    comp = np.zeros(256, dtype=np.uint8); comp[list(b"ACGT")] = list(b"TGCA")
    hdr = [int(np.argmax(im == 10)) + 1 for im in imgs]
    lens = [int((im.size - h) * width // (width + 1)) for im, h in zip(imgs, hdr)]

    def fetch(d, o, l):
        idx = o + np.arange(l, dtype=np.int64)
        return imgs[d][hdr[d] + idx + idx // width]

    # sample line starts without splitting a multi-hundred-MB buffer
    buf = np.frombuffer(text, dtype=np.uint8)
    rng = np.random.default_rng(0)
    starts = [0]
    for p in rng.integers(0, len(buf), how_many):
        nl = np.flatnonzero(buf[p:p + 4096] == 10)
        if len(nl):
            starts.append(int(p + nl[0] + 1))
    checked = 0
    for st in starts:
        if st >= len(buf):
            continue
        end = st + int(np.argmax(buf[st:st + 8192] == 10))
        fields = bytes(buf[st:end]).decode().split("\t")
        if len(fields) != 3:
            continue
        l, offs, strands = int(fields[0]), fields[1].split(","), fields[2].split(",")
        assert l >= 20 and len(offs) == len(imgs)
        ref = None
        if any(o != "" and int(o) + l > lens[d] for d, o in enumerate(offs)):
            continue                                   # a '+' MUM that runs through the document's '$' (reference quirk Q1): not checkable here
        for d, (o, sg) in enumerate(zip(offs, strands)):
            if o == "":
                continue
            seg = fetch(d, int(o), l)
            if sg == "-":
                seg = comp[seg[::-1]]
            if ref is None:
                ref = seg
            assert np.array_equal(seg, ref), f"occurrence mismatch in line at byte {st}"
        checked += 1
    return checked


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.smax, self.reasons = index, False, [], 0, set()

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.sm.append(float(out[0])); self.smax = max(self.smax, float(out[1]))
                for n, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.smax or None, "reasons": sorted(self.reasons)}


FLAGSETS = {      # reference CLI flags, keyword arguments of Engine.find
    "strict": ([], {}),
    "partial": (["-k", "-1"], {"num_distinct_docs": -1}),
    "mem": (["-f", "3", "-F", "100"], {"rare_freq": 3, "max_mem_freq": 100}),
}
REF_CAVEAT = ("reference sources compiled unmodified (oracle/Makefile: -O3 -march=x86-64-v3) against definitional stand-ins for its two un-vendored "
              "dependencies: sdsl (oracle/ref_shims/sdsl) and gsacak -> our SA-IS (oracle/sais.hpp, int64; slower than gsacak, which flatters the GPU/CPU ratio); "
              "the reference has no threads or OpenMP: 1 thread per process")


def sample_of(spec, L):
    """The same generator at L bases per document (the reference's time grows linearly with input size: see reference_modes)."""
    s = dict(spec); s["L"] = min(L, spec["L"])
    return s


def sample_name(sample):
    return f"{sample['N']}x{sample['L']} bp, generator of the workload with seed 1"


def reference_exe():
    exe = os.path.join(ROOT, "oracle", "_ref", "mumemto_exec")
    if os.path.exists(exe):
        return exe, "reference"
    exe = os.path.join(ROOT, "oracle", "_build", "mumemto_oracle")
    if not os.path.exists(exe):
        subprocess.run(["make", "port"], cwd=os.path.join(ROOT, "oracle"), check=True, capture_output=True)
    return exe, "port"


class SampleFiles:
    """A sample collection written to a scratch directory once; reference runs and the GPU parity check read the same images."""

    def __init__(self, sample):
        self.sample = sample
        self.imgs = make_collection(sample, seed=1)
        self.tmp = tempfile.TemporaryDirectory()
        self.dir = self.tmp.name
        self.paths = []
        for i, im in enumerate(self.imgs):
            p = os.path.join(self.dir, f"h{i}.fa"); im.tofile(p); self.paths.append(p)
        self.bases = sample["N"] * sample["L"]

    def groups(self, procs):
        """Partitions that share document 0 (reference README "mumemto merge"): [0] + every procs-th of the others."""
        procs = max(2, min(procs, len(self.paths) - 1))
        return [[0] + list(range(1 + g, len(self.paths), procs)) for g in range(procs)]

    def close(self):
        self.tmp.cleanup()


def run_reference(exe, files, flags, tag="ref"):
    """One single-process run of the reference on the sample; returns (seconds, output bytes)."""
    prefix = os.path.join(files.dir, tag)
    t0 = time.perf_counter()
    subprocess.run([exe, "-o", prefix, *flags, *files.paths], check=True, capture_output=True)
    secs = time.perf_counter() - t0
    out = b""
    for ext in (".mums", ".mems"):
        if os.path.exists(prefix + ext):
            with open(prefix + ext, "rb") as f:
                out = f.read()
    return secs, out


def run_reference_partitioned(exe, merge_exe, files, procs):
    """The reference's own way to use many cores (README "mumemto merge", mumemto/merge_mums.py:171-192): split the collection into
    partitions that share document 0, run one `mumemto_exec -M -n` PROCESS per partition concurrently, then `anchor_merge` the partition
    outputs.  The merged .mums is set-equal to a direct run on the whole collection.
    Returns (seconds, partitions as lists of document indices, merged .mums bytes, merged .athresh bytes)."""
    groups = files.groups(procs)
    d = files.dir
    t0 = time.perf_counter()
    running = [subprocess.Popen([exe, "-o", os.path.join(d, f"part{g}"), "-M", "-n", *[files.paths[i] for i in grp]],
                                stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for g, grp in enumerate(groups)]
    for pr in running:
        if pr.wait() != 0:
            raise RuntimeError("a reference partition run failed")
    subprocess.run([merge_exe, *[os.path.join(d, f"part{g}.mums") for g in range(len(groups))], "-o", os.path.join(d, "merged")],
                   check=True, capture_output=True)
    secs = time.perf_counter() - t0
    with open(os.path.join(d, "merged.mums"), "rb") as f:
        mums = f.read()
    with open(os.path.join(d, "merged.athresh"), "rb") as f:
        athresh = f.read()
    return secs, groups, mums, athresh


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def reference_modes_table(spec, flagset, sizes=(1_000_000, 4_000_000, 16_000_000), budget_s=200.0):
    """bases/s of the reference per mode and sample size (64 x {1, 4, 16} Mbp, linearity shown, the full-size
    figure is an extrapolation).  A size is skipped when the time predicted from the previous one does not fit what is left of the budget
    or the host lacks the memory (about 45 bytes per text symbol and process)."""
    exe, kind = reference_exe()
    merge_exe = os.path.join(ROOT, "oracle", "_ref", "anchor_merge")
    flags = FLAGSETS[flagset][0]
    cores = host_cores()
    can_partition = kind == "reference" and os.path.exists(merge_exe) and flagset == "strict" and spec["N"] >= 5 and spec["kind"] == "pangenome" and cores >= 2
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 32 << 30
    table = {"single_process": {"cores": 1, "sizes": {}}, "partition_processes": {"cores": 0, "sizes": {}}}
    t_start = time.perf_counter()
    last = {"single_process": None, "partition_processes": None}
    samples = {}
    try:
        for mode in ("partition_processes", "single_process"):          # the many-core mode first: it is the one the headline is compared with
            if mode == "partition_processes" and not can_partition:
                continue
            for L in sizes:
                if L > spec["L"]:
                    break
                sample = sample_of(spec, L)
                symbols = 2 * sample["N"] * (L + 1)
                left = budget_s - (time.perf_counter() - t_start)
                prev = last[mode]
                predicted = prev[1] * (L / prev[0]) if prev else 0.0
                need_mem = 45 * symbols * (1.3 if mode == "partition_processes" else 1.0)
                if predicted > left or need_mem > 0.8 * avail:
                    table[mode]["sizes"][str(L)] = {"skipped": f"predicted {predicted:.0f} s with {left:.0f} s of the budget left" if predicted > left else "host memory"}
                    continue
                if L not in samples:
                    samples[L] = SampleFiles(sample)
                files = samples[L]
                if mode == "single_process":
                    secs, _ = run_reference(exe, files, flags, tag=f"lin{L}")
                else:
                    secs, groups, _, _ = run_reference_partitioned(exe, merge_exe, files, cores)
                    table[mode]["cores"] = len(groups)
                last[mode] = (L, secs)
                table[mode]["sizes"][str(L)] = {"bases_per_s": files.bases / secs, "seconds": secs, "sample": sample_name(sample)}
    finally:
        for f in samples.values():
            f.close()
    return table


def reference_arm(args, spec, rank):
    """The reference's own CPU implementation of the path, timed on the host cores (rank 0 only; the other ranks exit).  Every step is
    one run of a bounded sample of the workload (config.workload names the sample that was timed, not the full size); the faster of
    the reference's two modes (one process; its partition-merge workflow with one process per core) is the line's value."""
    if rank != 0:
        return
    flagset = spec.get("flagset", "strict")
    exe, kind = reference_exe()
    merge_exe = os.path.join(ROOT, "oracle", "_ref", "anchor_merge")
    cores = host_cores()
    sample = sample_of(spec, int(os.environ.get("MMB_REF_SAMPLE_L", "1000000")))
    files = SampleFiles(sample)
    flags = FLAGSETS[flagset][0]
    partition_ok = kind == "reference" and os.path.exists(merge_exe) and flagset == "strict" and spec["N"] >= 5 and spec["kind"] == "pangenome" and cores >= 4
    procs = 0

    def step():
        nonlocal procs
        if partition_ok:
            secs, groups, _, _ = run_reference_partitioned(exe, merge_exe, files, cores)
            procs = len(groups)
            return secs
        return run_reference(exe, files, flags)[0]

    for _ in range(args.warmup):
        step()
    times = [step() for _ in range(max(1, args.steps))]
    files.close()
    t = float(np.mean(times))
    v = sample["N"] * sample["L"] / t
    used = procs if partition_ok else 1
    mode = (f"the reference's own parallel workflow: {procs} concurrent `mumemto_exec -M -n` processes (one partition each, sharing document 0) + anchor_merge"
            if partition_ok else "one process")
    modes = reference_modes_table(spec, flagset) if not os.environ.get("MMB_REF_NO_TABLE") else None
    line = {"impl": "reference", "metric": "input bases/sec end-to-end (FASTA->.mums)", "value": v, "unit": "bases/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"SAMPLE of {args.workload}: {sample_name(sample)} ({spec['desc']} is the full workload; the CPU figure for it is an extrapolation)",
                       "name": args.workload, "sample_bases": sample["N"] * sample["L"], "full_workload_bases": spec["N"] * spec["L"], "extrapolated": True,
                       "flags": flags},
            "cpu_baseline": {"value": v, "unit": "bases/s", "cores": used, "kind": kind, "sample": f"{sample_name(sample)}; {mode}; {REF_CAVEAT}", "host_cores": cores},
            "e2e": {"value": v, "unit": "bases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "reference_modes": modes, "extrapolated": True, "full_size_run": full_size_reference_run(spec)}
    print(json.dumps(line))


def snapshot_partition(eng, tdev):
    """Clone the engine's last MUM table + thresholds into torch tensors on the device (the engine's own buffers die at the next find)."""
    from mumemto_b200 import partition
    part = partition.local_partition(eng, tdev)
    return tuple(x.clone() for x in part)


def merge_on_device(eng, parts):
    """mmb_anchor_merge(inputs_on_device = 1) over cloned partition tables; returns (merged .mums bytes, merged .athresh bytes)."""
    args = [(p[0].data_ptr(), p[1].data_ptr(), p[2].data_ptr(), p[3].data_ptr(), int(p[0].shape[0]), int(p[1].shape[1])) for p in parts]
    eng.anchor_merge(args, int(parts[0][3].shape[0]), on_device=True)
    _, _, _, thr = eng.merge_result()
    return eng.merge_text(), thr.tobytes()


def cpu_baseline_and_parity(eng, tdev, spec, sample_L=1_000_000, procs=None):
    """Rank 0, N = 1, off the clock: the reference on a bounded sample of the workload (both modes), AND the byte-for-byte comparison of
    the GPU engine with the reference's output on that same sample -- direct run with the workload's flags, and (strict multi-MUMs)
    partitions with -M -n -> anchor_merge against the engine's partitions -> device-input merge.  Raises when any byte differs."""
    flagset = spec.get("flagset", "strict")
    flags, kw = FLAGSETS[flagset]
    exe, kind = reference_exe()
    merge_exe = os.path.join(ROOT, "oracle", "_ref", "anchor_merge")
    cores = procs or host_cores()
    sample = sample_of(spec, sample_L)
    files = SampleFiles(sample)
    parity = {}
    modes = {}
    # the reference's partition workflow first (a few seconds on >= 8 cores), then one process
    if kind == "reference" and os.path.exists(merge_exe) and flagset == "strict" and spec["kind"] == "pangenome" and spec["N"] >= 5 and cores >= 2:
        tp, groups, ref_mums, ref_athresh = run_reference_partitioned(exe, merge_exe, files, cores)
        modes["partition_processes"] = {"value": files.bases / tp, "cores": len(groups), "seconds": tp}
        parts = []
        for grp in groups:
            eng.load_fasta_bytes([files.imgs[i] for i in grp])
            eng.find(merge=True, anchor_merge=True, want_text=False)
            parts.append(snapshot_partition(eng, tdev))
        got_mums, got_athresh = merge_on_device(eng, parts)
        if got_mums != ref_mums or got_athresh != ref_athresh:
            raise AssertionError("parity_sample: device partition merge differs from the reference's anchor_merge output")
        parity["partitions_anchor_merge"] = {"partitions": len(groups), "mums_bytes": len(ref_mums), "athresh_bytes": len(ref_athresh), "result": "ok"}
    t1, ref_out = run_reference(exe, files, flags)
    modes["single_process"] = {"value": files.bases / t1, "cores": 1, "seconds": t1}
    eng.load_fasta_bytes(files.imgs)
    eng.find(want_text=True, **kw)
    got = eng.text()
    if got != ref_out:
        raise AssertionError(f"parity_sample: GPU output differs from the reference on {sample_name(sample)} with flags {flags}")
    parity["direct"] = {"flags": flags, "bytes": len(ref_out), "lines": ref_out.count(b"\n"), "result": "ok"}
    files.close()
    best = max(modes, key=lambda m: modes[m]["value"])
    full = full_size_reference_run(spec)
    return ({"full_size_run": full, "value": modes[best]["value"], "unit": "bases/s", "cores": modes[best]["cores"], "kind": kind, "seconds": modes[best]["seconds"],
             "sample": f"{sample_name(sample)}; mode {best}; {REF_CAVEAT}", "modes": modes, "extrapolated": True, "host_cores": cores},
            {"result": "ok", "sample": sample_name(sample), "checks": parity})


def cli_end_to_end(imgs, spec, flagset, expect_md5, runs=2):
    """Wall clock of `mumemto_exec` as a user runs it: process start, FASTA files on local disk (page cache), all the way to the closed
    .mums / .mems file.  The bench process keeps its own device buffers meanwhile (the executable needs the rest of the GPU's memory)."""
    import hashlib
    import shutil
    exe = os.path.join(ROOT, "mumemto_b200", "bin", "mumemto_exec")
    if not os.path.exists(exe):
        return {"error": "bin/mumemto_exec is not built"}
    d = tempfile.mkdtemp(prefix="mmb_cli_", dir=os.environ.get("MMB_CLI_DIR", "/tmp"))
    try:
        paths = []
        for i, im in enumerate(imgs):
            p = os.path.join(d, f"h{i}.fa"); np.asarray(im).tofile(p); paths.append(p)
        flags = FLAGSETS[flagset][0]
        times = []
        for _ in range(runs):
            t0 = time.perf_counter()
            r = subprocess.run([exe, "-o", os.path.join(d, "out"), *flags, *paths], capture_output=True, env=dict(os.environ, MMB_CLI_TIMING="1", MMB_PROFILE_HOST="1"))
            times.append(time.perf_counter() - t0)
            if r.returncode != 0:
                return {"error": r.stderr.decode(errors="replace")[-400:]}
        out = os.path.join(d, "out.mums" if flagset != "mem" else "out.mems")
        h = hashlib.md5()
        with open(out, "rb") as f:
            for blk in iter(lambda: f.read(1 << 24), b""):
                h.update(blk)
        best = min(times)
        return {"value": spec["N"] * spec["L"] / best, "unit": "bases/s", "seconds": best, "all_seconds": times, "output_bytes": os.path.getsize(out),
                "md5_matches_engine_output": h.hexdigest() == expect_md5, "stderr_tail": r.stderr.decode(errors="replace")[-1600:],
                "region": "wall clock of bin/mumemto_exec: process start, FASTA files in the page cache -> closed output file (pooled pread into pinned memory, H2D, path, D2H, write)"}
    finally:
        shutil.rmtree(d, ignore_errors=True)


def full_size_reference_run(spec):
    """The reference's one run on the FULL workload, when scripts/make_fullsize_golden.py has recorded it (single process, in the build
    container -- not on this host): the only same-config CPU figure there is; the sample-based figures above extrapolate."""
    name = next((k for k, v in WORKLOADS.items() if v is spec or (v["N"] == spec["N"] and v["L"] == spec["L"] and v.get("flagset") == spec.get("flagset") and v["kind"] == spec["kind"])), None)
    g = golden_md5(name, spec.get("flagset", "strict")) if name else None
    if not g:
        return None
    return {"seconds": g["reference_seconds"], "bases_per_s": g["input_bases"] / g["reference_seconds"], "peak_rss_gb": g.get("reference_peak_rss_gb"),
            "where": "build container (1 core; the reference has no threads), oracle/_ref/mumemto_exec on the full collection; output md5 = golden_full_size.reference_md5"}


def golden_md5(workload, flagset):
    """md5 of the reference's output on the FULL workload, when scripts/make_fullsize_golden.py has recorded it (tests/golden/md5_table.json)."""
    try:
        with open(os.path.join(ROOT, "tests", "golden", "md5_table.json")) as f:
            e = json.load(f).get(f"full:{workload}:{flagset}")
        return e
    except Exception:
        return None


# Algorithmic bytes per unit of the kernels that can dominate a step: what one launch must move if every operand
# were touched once.  unit: "suffix" = text symbol (|T|), "stream" = enumerated rank (after segment pruning), "dict" = dictionary
# position, "phrase" = phrase occurrence (parse symbol).
ROOF = {
    "expand_members": (29.0, "stream"),      # occurrence list 16 B read; SA 8 + LCP 4 + BWT 1 B written
    "detect_chunk": (5.0, "stream"),         # LCP 4 B + BWT 1 B read
    "detect_window": (4.0, "stream"),
    "segment_thresholds": (9.0, "suffix"),   # -M: 8 B list entry per suffix of a threshold-only segment + ~1 B of per-segment records
    "sort_groups_warp": (16.0, "listed suffix and round"),
    "sort_dgroups_warp": (33.0, "listed suffix and round"),   # sa 4 + 4, limit 4, 21 symbols of the dictionary
    "fasta_scatter_pass": (1.5, "suffix"),   # half a FASTA byte read, one symbol written
    "fasta_chunk_maps": (0.5, "suffix"),
    "trigger_masks": (1.125, "suffix"),      # 1 B read + 1/8 B written
    "phrase_hash": (1.12, "suffix"),         # every text byte once + 12 B per phrase
    "verify_groups": (2.0, "suffix"),        # the phrase and its representative
    "parse_lcp": (112.0, "phrase"),          # isa 4 + sa 4 + lcp 4 + the two differing phrases (~100 B)
    "dict_keys": (13.0, "dict"),             # 1 B + limit 4 read, key 8 written
    "dict_fill": (10.0, "dict"),
    "dict_info": (9.0, "dict"),
    "rank_words": (20.0, "dict"),            # key 8, sa 4, one gathered 4 B word, 4 B written
    "classify_segments": (5.0, "dict"),
    "member_flags": (24.0, "dict"),
    "list_dict_groups": (16.0, "dict"),
    "first_flags": (12.0, "dict"),
    "cub_radix_sort": (None, None),          # library kernels: time only
    "cub_scan": (None, None),
}


PEAK_GBS = 6650.0
try:
    with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as _f:
        PEAK_GBS = float(json.load(_f).get("hbm_gbs", PEAK_GBS))
except Exception:
    pass


def ncu_traffic(kernel_name, st):
    """DRAM traffic of a kernel per step, from its `ncu --set full` capture: profiles/ncu/traffic.json holds dram__bytes_read.sum +
    dram__bytes_write.sum of one captured launch divided by the units that launch processed; scaled here by this run's units.  A
    constant of the capture (taken at 64 x 16 Mbp unless noted), not measured in this run -- the source string says so."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu", "traffic.json")) as f:
            table = json.load(f)
    except Exception:
        return None
    for key, tr in table.items():
        if key in kernel_name:
            u = roof_units(tr["unit"], st)
            if not u:
                return None
            return {"ncu_dram_bytes_per_unit": tr["dram_bytes_per_unit"], "ncu_dram_bytes_per_step": tr["dram_bytes_per_unit"] * u,
                    "ncu_traffic_source": f"profiles/ncu/traffic.json: {tr['dram_bytes_per_unit']} B per {tr['unit']} unit from {tr['capture']} "
                                          f"x this run's units (a constant of the capture, not measured in this run)"}
    return None


def roof_units(unit, st):
    return {"suffix": float(st["text_symbols"]), "stream": float(st.get("stream_suffixes") or st["text_symbols"]), "dict": float(st["pfp_dict_bytes"]),
            "phrase": float(st["pfp_phrases"])}.get(unit)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("MMB_WORKLOAD", "c2"))
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU baseline AND the parity-sample check (profiling runs only)")
    ap.add_argument("--route", default=os.environ.get("MMB_ROUTE", "auto"), choices=["auto", "direct", "pfp"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--sharded", action="store_true", help="N > 1: ONE collection (the N = 1 workload) on N GPUs, every rank reports its slice of the suffix "
                    "order (SURVEY 8e(2)); strong scaling.  Without it N > 1 runs the partition-merge workload (one partition per GPU)")
    ap.add_argument("--merge-mode", action="store_true", help="N = 1 only: run the partition flags (-M -n) without an exchange, to time that mode")
    ap.add_argument("--tile-ranks", type=int, default=0, help="ranks per stream tile on the PFP route (0 = engine default)")
    ap.add_argument("--roofline-kernel", default=None, help="kernel (name substring) whose launches are timed with CUDA events for the roofline line")
    ap.add_argument("--kernel-table-steps", type=int, default=2, help="extra steps after the timed region with every launch bracketed by events (per-kernel table)")
    ap.add_argument("--kernel-table-rows", type=int, default=16)
    ap.add_argument("--cli-e2e", action="store_true", help="also time bin/mumemto_exec on the workload written to local disk (wall clock of the whole process)")
    ap.add_argument("--dump-output", default=None, help="write the .mums/.mems bytes of the last end-to-end step to this file")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    sharded = world > 1 and args.sharded
    if world > 1 and args.workload == "c2" and not sharded:
        args.workload = "c5"            # several GPUs: the partition-merge configuration, one 32-genome partition per GPU
    spec = WORKLOADS[args.workload]
    flagset = spec.get("flagset", "strict")
    find_kw = dict(FLAGSETS[flagset][1])

    if args.impl == "reference":
        reference_arm(args, spec, rank)
        return

    import torch
    import torch.distributed as dist
    from mumemto_b200 import Engine

    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # NCCL's banner / debug lines must not mix with the JSON line on stdout
        # NCCL prints "NCCL version ..." straight to fd 1 when its first communicator comes up: point fd 1 at stderr until that has happened
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    stream = torch.cuda.current_stream()
    eng = Engine(device=local, stream=stream.cuda_stream)

    # ---- synthetic input: N = 1 -> the named collection; N > 1 -> one partition per rank sharing the anchor document
    multi = world > 1 and not sharded           # partition-merge mode; `sharded`: every rank holds the whole N = 1 collection
    cache = os.environ.get("MMB_CACHE_DIR")      # optional: keep the generated FASTA images between runs on the same box
    cache_file = os.path.join(cache, f"{args.workload}_r{rank}_w{world}.npy") if cache else None
    imgs = None
    if cache_file and os.path.exists(cache_file):
        data = np.load(cache_file, mmap_mode="r")
        sz = np.load(cache_file + ".sizes.npy")
        at0 = np.concatenate([[0], np.cumsum(sz)[:-1]])
        imgs = [data[o:o + n] for o, n in zip(at0, sz)]
    imgs = imgs or make_collection(spec, seed=1 + 1000 * (rank if multi else 0), anchor_seed=(12345 if (multi and spec["kind"] == "pangenome") else None),
                           first_index=0) if not (multi and spec["kind"] == "independent") else make_collection(spec, seed=1)
    if multi and spec["kind"] == "independent":
        # partitions of the Appendix-A collection cannot share an anchor unless it is the same h0: rank r keeps h0 + its own variants
        rng = np.random.default_rng(1)
        base = rng.integers(0, 4, spec["L"])
        def variant(seed):
            r2 = np.random.default_rng(seed); s = base.copy(); m = r2.random(spec["L"]) < spec["snp"]
            s[m] = (s[m] + r2.integers(1, 4, m.sum())) % 4
            return s.astype(np.uint8)
        imgs = [fasta_image("h0", variant(7))] + [fasta_image(f"h{rank * 100 + i}", variant(1000 * rank + i + 11)) for i in range(1, spec["N"])]
    if cache_file and not os.path.exists(cache_file):
        os.makedirs(cache, exist_ok=True)
        np.save(cache_file + ".sizes.npy", np.array([im.size for im in imgs], dtype=np.int64))
        np.save(cache_file, np.concatenate(imgs))
    bases_rank = spec["N"] * spec["L"]
    new_bases_rank = bases_rank if not multi else (spec["N"] - 1) * spec["L"]      # the anchor is counted once
    total_bases = bases_rank if not multi else spec["L"] + world * new_bases_rank

    sizes = [int(im.size) for im in imgs]
    offsets, at = [], 0
    for s in sizes:
        offsets.append(at); at += (s + 15) & ~15
    host = torch.empty(at + 32, dtype=torch.uint8).pin_memory()
    hv = host.numpy()
    for o, im in zip(offsets, imgs):
        hv[o:o + im.size] = im
    dev = host.cuda(non_blocking=False)
    pinned_imgs = [hv[o:o + s] for o, s in zip(offsets, sizes)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")          # > 126 MB L2
    merge_flags = multi or args.merge_mode
    opts = dict(merge=merge_flags, anchor_merge=merge_flags, route=args.route, tile_ranks=args.tile_ranks, **find_kw)
    if sharded:
        opts["shard"] = (rank, world)

    from mumemto_b200 import partition
    tdev = torch.device("cuda", local)
    exchange_ev = []       # (start, gathered, merged) events of the exchange step inside timed steps (N > 1)

    def exchange(record):
        """The exchange step of the path (N > 1): ONE all-gather of the partition tables, then the merge kernels on every rank."""
        if record:
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record(stream)
        parts = partition.gather_partitions(partition.local_partition(eng, tdev), dist, world)
        if record:
            e1.record(stream)
        partition.merge_gathered(eng, parts, fetch=False)
        if record:
            e2.record(stream); exchange_ev.append((e0, e1, e2))

    def step_device(record=False):
        eng.load_fasta_device(dev.data_ptr(), offsets, sizes)
        st = eng.find(want_text=True, **opts)
        if multi:
            exchange(record)
        if sharded:         # the final gather of the output slices (the only exchange of this mode), inside the timed step
            slices_out[0], slices_out[1] = partition.gather_slices(eng, dist, world, tdev)
        return st

    slices_out = [None, 0]
    out_host = None
    merged_host = None

    def step_e2e():
        nonlocal out_host, merged_host
        eng.load_fasta_bytes(pinned_imgs)
        eng.find(want_text=True, **opts)
        if sharded:         # every rank uploaded the whole collection; the slices are gathered on the devices and rank 0 reads them back
            allb, n = partition.gather_slices(eng, dist, world, tdev)
            if rank == 0:
                if out_host is None or out_host.numel() < n:
                    out_host = torch.empty(max(n + n // 8, 1), dtype=torch.uint8).pin_memory()
                out_host[:n].copy_(allb, non_blocking=False)
            return n
        n = eng.lib.mmb_text_size(eng.h)
        if out_host is None or out_host.numel() < n:
            out_host = torch.empty(max(n, 1), dtype=torch.uint8).pin_memory()
        eng.copy_text_into(out_host.numpy())
        if multi:       # partition .mums above, then exchange + merge; rank 0 brings the merged .mums back to the host (pinned buffer)
            exchange(False)
            if rank == 0:
                mn = eng.merge_text_size()
                if merged_host is None or merged_host.numel() < mn:
                    merged_host = torch.empty(max(mn + mn // 8, 1), dtype=torch.uint8).pin_memory()
                n += eng.merge_text_into(merged_host.numpy())
        return n

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    min_warm = int(os.environ.get("MMB_MIN_WARMUP", "3"))      # profiling runs only; a bench line always has >= 3 warm-up steps
    n_warm = max(args.warmup, min_warm, 1)
    eng.profile_kernel("*")                                     # the warm-up steps also tell which of OUR kernels dominates the step
    for _ in range(n_warm):
        flush.zero_(); st_warm = step_device()
    torch.cuda.synchronize()
    warm_table = eng.profile_table()
    own = sorted(((v[0], k) for k, v in warm_table.items() if not k.startswith("cub_")), reverse=True)
    roof_kernel = args.roofline_kernel or (own[0][1][:-2] if own and own[0][1].endswith("_k") else (own[0][1] if own else "expand_members"))
    eng.profile_kernel(roof_kernel)
    sampler = ClockSampler(local); sampler.start()
    barrier()
    launches0 = eng.stats()["kernel_launches"]
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    stage = {k: 0.0 for k in ("ms_text", "ms_sa", "ms_lcp", "ms_detect", "ms_emit")}
    st = None
    t_wall0 = time.perf_counter()
    per_step_stages = []
    for a, b in ev:
        flush.zero_()
        a.record(stream)
        st = step_device(record=True)
        b.record(stream)
        for k in stage:
            stage[k] += st[k]
        per_step_stages.append({k: round(st[k], 2) for k in stage})
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = eng.stats()["kernel_launches"] - launches0
    roof_ms, roof_launches = eng.profile_result()
    eng.profile_kernel(None)
    step_ms = [a.elapsed_time(b) for a, b in ev]
    ms = float(np.mean(step_ms))
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = total_bases / (ms_max * 1e-3)
    sampler.stop_flag = True; sampler.join(timeout=2)
    exchange_ms = None
    if exchange_ev:
        exchange_ms = {"allgather_ms": float(np.mean([a.elapsed_time(b) for a, b, _ in exchange_ev])),
                       "merge_ms": float(np.mean([b.elapsed_time(c) for _, b, c in exchange_ev]))}

    # ---- per-kernel table: a few extra steps with EVERY launch bracketed by events (outside the timed region: the brackets add no work
    # but the table is not part of the headline number)
    kernel_table = None
    if args.kernel_table_steps > 0:
        eng.profile_kernel("*")
        for _ in range(args.kernel_table_steps):
            flush.zero_(); step_device()
        torch.cuda.synchronize()
        tab = eng.profile_table()
        eng.profile_kernel(None)
        tot = sum(v[0] for v in tab.values()) or 1.0
        top = sorted(tab.items(), key=lambda kv: -kv[1][0])[:args.kernel_table_rows]
        kernel_table = []
        for name, (kms, kn) in top:
            per_step = kms / args.kernel_table_steps
            entry = {"kernel": name, "ms_per_step": round(per_step, 3), "launches_per_step": kn / args.kernel_table_steps, "share_of_kernel_time": round(kms / tot, 4)}
            for key, (bpu, unit) in ROOF.items():
                u = roof_units(unit, st) if bpu else None
                if key in name and bpu and u:
                    gbs = bpu * u / (per_step * 1e-3) / 1e9
                    entry.update({"algorithmic_bytes_per_unit": bpu, "unit": unit, "units_per_step": u, "achieved_gbs": round(gbs, 1),
                                  "frac_of_measured_peak": round(gbs / PEAK_GBS, 4)})
            tr = ncu_traffic(name, st)
            if tr:
                entry.update(tr)
            kernel_table.append(entry)

    # ---- end to end through the host-buffer call
    for _ in range(0 if args.no_e2e else 2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    out_bytes = 0
    e2e_stage = {k: 0.0 for k in stage}
    for _ in range(1 if args.no_e2e else args.steps):
        out_bytes = step_e2e()
        s2 = eng.stats()
        for k in e2e_stage:
            e2e_stage[k] += s2[k] / args.steps
    barrier()
    e2e_s = (time.perf_counter() - t0) / (1 if args.no_e2e else args.steps)
    te = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = total_bases / float(te.item())

    total_matches = int(st["num_matches"])
    if sharded:
        tm = torch.tensor([total_matches], device="cuda", dtype=torch.int64)
        dist.all_reduce(tm)
        total_matches = int(tm.item())

    # ---- checks off the clock (rank 0): sampled lines against the sequences, md5 of the whole output against the reference's run on the
    # full workload when that is on record, and the byte-for-byte parity sample against the reference executed here
    verified = 0
    output_md5 = None
    golden_check = None
    if rank == 0 and not multi and out_host is not None:
        out_np = out_host.numpy()[:out_bytes]
        if eng.mum_mode:
            verified = verify_sample(imgs, bytes(out_np))
        import hashlib
        output_md5 = hashlib.md5(out_np).hexdigest()
        if args.dump_output:
            out_np.tofile(args.dump_output)
        g = golden_md5(args.workload, flagset) if not args.merge_mode else None
        if g:
            ext = ".mums" if eng.mum_mode else ".mems"
            golden_check = {"reference_md5": g["md5"].get(ext), "reference_lines": g["lines"].get(ext), "match": g["md5"].get(ext) == output_md5}
            if not golden_check["match"]:
                raise AssertionError(f"output md5 {output_md5} differs from the reference's {g['md5'].get(ext)} on the full workload {args.workload}")

    # ---- partition merge (N > 1): the same exchange once more with wall-clock brackets and the table sizes
    merge_info = None
    if multi:
        merge_info = partition_merge(eng, dist, world, rank, spec)
        if exchange_ms:
            merge_info.update(exchange_ms)

    cpu_base = parity = None
    if rank == 0 and not args.no_cpu_baseline:
        # at N > 1 rank 0 runs the same parity sample (the partition-merge check is the one that covers the multi-GPU path)
        cpu_base, parity = cpu_baseline_and_parity(eng, tdev, spec if not multi else dict(WORKLOADS["c2"], L=spec["L"]))
    if world > 1:
        dist.barrier()

    # ---- the real end to end: the drop-in executable, process start with FASTA files on local disk -> closed output file.
    # Last: this process gives its device memory back first (the executable is a separate process on the same GPU).
    e2e_cli = None
    if rank == 0 and not multi and args.cli_e2e and not args.merge_mode:
        eng.close()
        del dev, flush
        torch.cuda.empty_cache()
        e2e_cli = cli_end_to_end(imgs, spec, flagset, output_md5)

    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        peak = peaks.get("hbm_gbs", 6650.0)
        algo_bytes = 60.0 * bases_rank                       # 30 B per suffix, 2 suffixes per input base
        achieved = algo_bytes / (ms * 1e-3) / 1e9
        roof_entry = next(((k, v) for k, v in ROOF.items() if k in roof_kernel), (roof_kernel, (None, None)))
        roof_unit = roof_entry[1][1]
        units = (roof_units(roof_unit, st) or float(st["text_symbols"])) * args.steps       # units the kernel processed over the timed region
        if roof_unit == "listed suffix and round":
            units = float(st["sa_sorted_elements"]) * args.steps
        # DRAM bytes per suffix of that kernel from its ncu --set full capture (dram__bytes_read.sum + dram__bytes_write.sum of one launch
        # divided by the suffixes that launch produced); the capture and its summary are under profiles/ncu/
        traffic = None; traffic_src = None
        tr = ncu_traffic(roof_kernel, st)
        if tr and roof_launches:
            traffic = tr["ncu_dram_bytes_per_step"] * args.steps / roof_launches
            traffic_src = tr["ncu_traffic_source"]
        bpu = roof_entry[1][0] or 0.0
        k_bytes = bpu * units
        k_achieved = (k_bytes / (roof_ms * 1e-3) / 1e9) if roof_ms > 0 else None
        line = {"metric": "input bases/sec end-to-end (FASTA->.mums)", "value": value, "unit": "bases/s", "n_gpus": world, "steps": args.steps,
                "warmup": n_warm, "ms_per_step": ms_max, "higher_is_better": True, "scaling": "strong" if sharded else "weak", "vs_baseline": None, "dtype": "u8",
                "data": "synthetic", "config": {"workload": spec["desc"], "name": args.workload, "flags": FLAGSETS[flagset][0] + (["-M", "-n"] if merge_flags else []),
                                                "l2": "flushed between steps (256 MiB write)",
                                                "partitioning": (f"1 collection on every GPU, {world} slices of the suffix order (1 per GPU), output slices gathered in order [SURVEY 8e(2)]" if sharded
                                                                 else "1 collection" if not multi else f"{world} partitions sharing anchor h0, -M -n, 1 per GPU")},
                "value_region": "device: FASTA images resident in HBM -> formatted output bytes in HBM (CUDA events); e2e adds the host<->device copies; e2e_cli (when present) is process start -> closed file",
                "e2e": {"value": e2e_value, "unit": "bases/s", "h2d_bytes_per_step": int(sum(sizes)) * (world if sharded else 1), "d2h_bytes_per_step": int(out_bytes)},
                "gpu_launches": int(launches),
                "roofline": {"bound": "hbm", "kernel": roof_kernel, "achieved": k_achieved, "peak": peak, "unit": "GB/s",
                             "frac": (k_achieved / peak) if k_achieved else None, "traffic": traffic, "traffic_source": traffic_src,
                             "achieved_per_launch_bytes": (k_bytes / roof_launches) if roof_launches else None,
                             "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback", "kernel_ms_per_step": roof_ms / args.steps,
                             "kernel_launches_per_step": roof_launches / args.steps, "kernel_share_of_step": (roof_ms / args.steps) / ms if ms else None,
                             "algorithmic_bytes_per_unit": bpu, "unit": roof_unit, "units_per_step": units / args.steps,
                             "whole_path": {"achieved": achieved, "frac": achieved / peak, "algorithmic_bytes_per_base": 60.0},
                             "kernels": kernel_table},
                "stages_ms": {k: v / args.steps for k, v in stage.items()}, "step_ms": step_ms, "per_step_stages": per_step_stages, "e2e_stages_ms": e2e_stage,
                "matches": int(total_matches), "verified_sample_lines": int(verified), "output_md5": output_md5, "golden_full_size": golden_check,
                "parity_sample": parity["result"] if parity else None, "parity_sample_detail": parity,
                "sa_rounds": int(st["sa_rounds"]), "sa_sorted_elements": int(st["sa_sorted_elements"]),
                "text_symbols": int(st["text_symbols"]), "wall_s_timed_region": t_wall, "clocks": sampler.summary(),
                "route": int(st["route"]), "pruned": int(st.get("pruned", 0)), "stream_suffixes": int(st.get("stream_suffixes", 0)),
                "threshold_segments": int(st.get("threshold_segments", 0)), "pfp": {k: int(st[k]) for k in ("pfp_phrases", "pfp_distinct_phrases", "pfp_dict_bytes", "pfp_tiles")}}
        if e2e_cli:
            line["e2e_cli"] = e2e_cli
        if merge_info:
            line["partition_merge"] = merge_info
        if cpu_base and not multi:                          # the CPU figure is reported at N = 1 only; at N > 1 the run above serves the parity check
            line["cpu_baseline"] = cpu_base
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def partition_merge(eng, dist, world, rank, spec):
    """One all-gather of every rank's partition table (device buffers, NCCL) and the merge kernels on every rank."""
    import torch
    from mumemto_b200 import partition
    dev = torch.device("cuda", torch.cuda.current_device())
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    parts = partition.gather_partitions(partition.local_partition(eng, dev), dist, world)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    partition.merge_gathered(eng, parts, fetch=False)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    return {"in_timed_step": True, "allgather_s": t1 - t0, "merge_s": t2 - t1, "merged_mums": int(eng.lib.mmb_merge_num_mums(eng.h)),
            "merged_columns": int(eng.lib.mmb_merge_num_docs(eng.h)), "partition_mums": [int(p[4]) for p in parts],
            "gathered_bytes": int(sum(p[0].numel() * 4 + p[1].numel() * 8 + p[2].numel() + p[3].numel() * 2 for p in parts))}


if __name__ == "__main__":
    main()
