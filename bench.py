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


def bounded_sample(spec, budget_bases=8_000_000):
    """A collection of the same generator small enough for the single-threaded reference: ~5-10 s of CPU per run."""
    sample = dict(spec)
    note = "whole workload"
    if spec["N"] * spec["L"] > budget_bases:
        sample["L"] = max(10_000, budget_bases // spec["N"])
        note = f"{spec['N']}x{sample['L']} bp collection from the same generator (the reference's time grows linearly with input size)"
    return sample, note


def reference_exe():
    exe = os.path.join(ROOT, "oracle", "_ref", "mumemto_exec")
    if os.path.exists(exe):
        return exe, "reference"
    exe = os.path.join(ROOT, "oracle", "_build", "mumemto_oracle")
    if not os.path.exists(exe):
        subprocess.run(["make", "port"], cwd=os.path.join(ROOT, "oracle"), check=True, capture_output=True)
    return exe, "port"


def run_reference(exe, sample, repeats, warmup):
    imgs = make_collection(sample, seed=1)
    times = []
    with tempfile.TemporaryDirectory() as d:
        paths = []
        for i, im in enumerate(imgs):
            p = os.path.join(d, f"h{i}.fa"); im.tofile(p); paths.append(p)
        for it in range(warmup + repeats):
            t0 = time.perf_counter()
            subprocess.run([exe, "-o", os.path.join(d, "ref"), *paths], check=True, capture_output=True)
            if it >= warmup:
                times.append(time.perf_counter() - t0)
    return float(np.mean(times))


def run_reference_partitioned(exe, merge_exe, sample, procs, repeats, warmup):
    """The reference's own way to use many cores (README "mumemto merge", mumemto/merge_mums.py:171-192): split the collection into
    partitions that share document 0, run one `mumemto_exec -M -n` PROCESS per partition concurrently, then `anchor_merge` the partition
    outputs.  The merged .mums is set-equal to a direct run on the whole collection.  Returns (seconds per run, partitions)."""
    imgs = make_collection(sample, seed=1)
    N = len(imgs)
    procs = max(2, min(procs, N - 1))
    times = []
    with tempfile.TemporaryDirectory() as d:
        paths = []
        for i, im in enumerate(imgs):
            p = os.path.join(d, f"h{i}.fa"); im.tofile(p); paths.append(p)
        groups = [[paths[0]] + paths[1 + g::procs] for g in range(procs)]
        for it in range(warmup + repeats):
            t0 = time.perf_counter()
            running = [subprocess.Popen([exe, "-o", os.path.join(d, f"part{g}"), "-M", "-n", *grp], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
                       for g, grp in enumerate(groups)]
            for pr in running:
                if pr.wait() != 0:
                    raise RuntimeError("a reference partition run failed")
            subprocess.run([merge_exe, *[os.path.join(d, f"part{g}.mums") for g in range(procs)], "-o", os.path.join(d, "merged")],
                           check=True, capture_output=True)
            if it >= warmup:
                times.append(time.perf_counter() - t0)
    return float(np.mean(times)), procs


def reference_best(spec, repeats, warmup):
    """Both ways the reference can run the path on this host: one process (it has no threads), and its partition-merge workflow with
    one process per core.  Returns the faster one as (bases/s, seconds per run, cores, kind, sample note, both modes)."""
    exe, kind = reference_exe()
    sample, note = bounded_sample(spec)
    t1 = run_reference(exe, sample, repeats, warmup)
    modes = {"single_process": {"value": sample["N"] * sample["L"] / t1, "cores": 1, "seconds": t1, "sample": note}}
    best = (modes["single_process"]["value"], t1, 1, kind,
            note + "; reference sources compiled unmodified with sdsl/gsacak stand-ins (oracle/ref_shims), 1 thread: the reference has no threads or OpenMP")
    merge_exe = os.path.join(ROOT, "oracle", "_ref", "anchor_merge")
    cores = os.cpu_count() or 1
    if kind == "reference" and os.path.exists(merge_exe) and cores >= 4 and spec["N"] >= 5 and spec["kind"] == "pangenome":
        big, bnote = bounded_sample(spec, budget_bases=8_000_000 * min(cores, spec["N"] - 1) // 2)
        try:
            tp, procs = run_reference_partitioned(exe, merge_exe, big, cores, repeats, warmup)
            vp = big["N"] * big["L"] / tp
            modes["partition_processes"] = {"value": vp, "cores": procs, "seconds": tp, "sample": bnote}
            if vp > best[0]:
                best = (vp, tp, procs, kind, bnote + f"; the reference's own parallel workflow: {procs} concurrent mumemto_exec -M -n processes (one partition each, "
                        "sharing document 0) + anchor_merge; reference sources compiled unmodified with sdsl/gsacak stand-ins")
        except Exception as ex:      # the single-process number stands
            modes["partition_processes"] = {"error": str(ex)[:200]}
    return best + (modes,)


def reference_arm(args, spec, rank):
    """The reference's own CPU implementation of the path, timed on the host cores (rank 0 only; the other ranks exit)."""
    if rank != 0:
        return
    v, t, cores, kind, note, modes = reference_best(spec, args.steps, min(args.warmup, 1))
    line = {"impl": "reference", "metric": "input bases/sec end-to-end (FASTA->.mums)", "value": v, "unit": "bases/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic", "config": {"workload": spec["desc"], "name": args.workload},
            "cpu_baseline": {"value": v, "unit": "bases/s", "cores": cores, "kind": kind, "sample": note},
            "e2e": {"value": v, "unit": "bases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "reference_modes": modes}
    print(json.dumps(line))


def cpu_baseline_sample(spec):
    """Bounded CPU run beside the GPU number (rank 0, N = 1): the reference build on (a sample of) the workload, in the faster of its
    two modes (one process; one process per core through its partition-merge workflow)."""
    v, t, cores, kind, note, modes = reference_best(spec, 1, 0)
    return {"value": v, "unit": "bases/s", "cores": cores, "kind": kind, "sample": note, "seconds": t, "modes": modes}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("MMB_WORKLOAD", "c2"))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--route", default=os.environ.get("MMB_ROUTE", "auto"), choices=["auto", "direct", "pfp"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--merge-mode", action="store_true", help="N = 1 only: run the partition flags (-M -n) without an exchange, to time that mode")
    ap.add_argument("--tile-ranks", type=int, default=0, help="ranks per stream tile on the PFP route (0 = engine default)")
    ap.add_argument("--roofline-kernel", default=None, help="kernel (name substring) whose launches are timed with CUDA events for the roofline line")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and args.workload == "c2":
        args.workload = "c5"            # several GPUs: the partition-merge configuration, one 32-genome partition per GPU
    spec = WORKLOADS[args.workload]

    if args.impl == "reference":
        reference_arm(args, spec, rank)
        return

    import torch
    import torch.distributed as dist
    from mumemto_b200 import Engine

    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # NCCL's banner / debug lines must not mix with the JSON line on stdout
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.current_stream()
    eng = Engine(device=local, stream=stream.cuda_stream)

    # ---- synthetic input: N = 1 -> the named collection; N > 1 -> one partition per rank sharing the anchor document
    multi = world > 1
    cache = os.environ.get("MMB_CACHE_DIR")      # optional: keep the generated FASTA images between runs on the same box
    cache_file = os.path.join(cache, f"{args.workload}_r{rank}_w{world}.npy") if cache else None
    imgs = None
    if cache_file and os.path.exists(cache_file):
        data = np.load(cache_file, mmap_mode="r")
        sz = np.load(cache_file + ".sizes.npy")
        at0 = np.concatenate([[0], np.cumsum(sz)[:-1]])
        imgs = [data[o:o + n] for o, n in zip(at0, sz)]
    imgs = imgs or make_collection(spec, seed=1 + 1000 * rank, anchor_seed=(12345 if (multi and spec["kind"] == "pangenome") else None),
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
    opts = dict(merge=multi or args.merge_mode, anchor_merge=multi or args.merge_mode, route=args.route, tile_ranks=args.tile_ranks)

    from mumemto_b200 import partition
    tdev = torch.device("cuda", local)

    def step_device():
        eng.load_fasta_device(dev.data_ptr(), offsets, sizes)
        st = eng.find(want_text=True, **opts)
        if multi:       # the exchange step of the path: one all-gather of the partition tables, then the merge kernels on every rank
            partition.partition_merge(eng, dist, world, tdev, fetch=False)
        return st

    out_host = None
    merged_host = None

    def step_e2e():
        nonlocal out_host, merged_host
        eng.load_fasta_bytes(pinned_imgs)
        eng.find(want_text=True, **opts)
        n = eng.lib.mmb_text_size(eng.h)
        if out_host is None or out_host.numel() < n:
            out_host = torch.empty(max(n, 1), dtype=torch.uint8).pin_memory()
        eng.copy_text_into(out_host.numpy())
        if multi:       # partition .mums above, then exchange + merge; rank 0 brings the merged .mums back to the host (pinned buffer)
            partition.partition_merge(eng, dist, world, tdev, fetch=False)
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
    for _ in range(max(args.warmup, min_warm, 1)):
        flush.zero_(); st_warm = step_device()
    # dominant kernel per route (profiles/README.md): algorithmic bytes per suffix it must move
    ROOF = {"expand_members": 29.0,      # occurrence list 16 B read; SA 8 + LCP 4 + BWT 1 B written
            "detect_window": 4.0,        # the LCP tile; SA/BWT only for surviving candidates
            "sort_groups_warp": 16.0,    # per listed suffix and round: sa 4+4, isa 4 (random), nh 4
            "fasta_scatter_pass": 1.5}   # per suffix: half a FASTA byte read, one symbol written
    roof_kernel = args.roofline_kernel or ("expand_members" if st_warm["route"] == 2 else "sort_groups_warp")
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
        st = step_device()
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
    sampler.stop_flag = True; sampler.join(timeout=2)
    verified = verify_sample(imgs, bytes(out_host.numpy()[:out_bytes]) if (out_host is not None and rank == 0 and not multi) else b"")

    # ---- partition merge (N > 1): one NCCL all-gather of the partition tables, then the merge kernels on every rank
    merge_info = None
    if multi:
        merge_info = partition_merge(eng, dist, world, rank, spec)

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
        units = float(st["text_symbols"]) * args.steps       # suffixes the kernel processed over the timed region
        if roof_kernel == "sort_groups_warp":
            units = float(st["text_symbols"]) * max(1, int(st["sa_rounds"])) * args.steps      # upper bound: every suffix listed in every round
        # DRAM bytes per suffix of that kernel from its ncu --set full capture (dram__bytes_read.sum + dram__bytes_write.sum of one launch
        # divided by the suffixes that launch produced); the capture and its summary are under profiles/ncu/
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "ncu", "traffic.json")) as f:
                tr = json.load(f).get(roof_kernel)
            if tr and roof_launches:
                traffic = tr["dram_bytes_per_suffix"] * float(st["text_symbols"]) * args.steps / roof_launches
        except Exception:
            pass
        k_bytes = ROOF.get(roof_kernel, 0.0) * units
        k_achieved = (k_bytes / (roof_ms * 1e-3) / 1e9) if roof_ms > 0 else None
        line = {"metric": "input bases/sec end-to-end (FASTA->.mums)", "value": value, "unit": "bases/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_max, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
                "data": "synthetic", "config": {"workload": spec["desc"], "name": args.workload, "l2": "flushed between steps (256 MiB write)",
                                                "partitioning": "1 collection" if not multi else f"{world} partitions sharing anchor h0, -M -n, 1 per GPU"},
                "e2e": {"value": e2e_value, "unit": "bases/s", "h2d_bytes_per_step": int(sum(sizes)), "d2h_bytes_per_step": int(out_bytes)},
                "gpu_launches": int(launches),
                "roofline": {"bound": "hbm", "kernel": roof_kernel, "achieved": k_achieved, "peak": peak, "unit": "GB/s",
                             "frac": (k_achieved / peak) if k_achieved else None, "traffic": traffic, "achieved_per_launch_bytes": (k_bytes / roof_launches) if roof_launches else None,
                             "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback", "kernel_ms_per_step": roof_ms / args.steps,
                             "kernel_launches_per_step": roof_launches / args.steps, "kernel_share_of_step": (roof_ms / args.steps) / ms if ms else None,
                             "algorithmic_bytes_per_suffix": ROOF.get(roof_kernel),
                             "whole_path": {"achieved": achieved, "frac": achieved / peak, "algorithmic_bytes_per_base": 60.0}},
                "stages_ms": {k: v / args.steps for k, v in stage.items()}, "step_ms": step_ms, "per_step_stages": per_step_stages, "e2e_stages_ms": e2e_stage,
                "matches": int(st["num_matches"]), "verified_sample_lines": int(verified), "sa_rounds": int(st["sa_rounds"]), "sa_sorted_elements": int(st["sa_sorted_elements"]),
                "text_symbols": int(st["text_symbols"]), "wall_s_timed_region": t_wall, "clocks": sampler.summary(),
                "route": int(st["route"]), "pfp": {k: int(st[k]) for k in ("pfp_phrases", "pfp_distinct_phrases", "pfp_dict_bytes", "pfp_tiles")}}
        if merge_info:
            line["partition_merge"] = merge_info
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_sample(spec)
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
    partition.merge_gathered(eng, parts)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    return {"in_timed_step": True, "allgather_s": t1 - t0, "merge_s": t2 - t1, "merged_mums": int(eng.lib.mmb_merge_num_mums(eng.h)),
            "merged_columns": int(eng.lib.mmb_merge_num_docs(eng.h)), "partition_mums": [int(p[4]) for p in parts],
            "gathered_bytes": int(sum(p[0].numel() * 4 + p[1].numel() * 8 + p[2].numel() + p[3].numel() * 2 for p in parts))}


if __name__ == "__main__":
    main()
