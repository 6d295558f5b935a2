"""Host-side mirror of the reference's operator interface for the FASTA -> .mums/.mems path.

* ``mum`` / ``mem``: same argument names, meaning and error behaviour as ``mumemto.mum`` / ``mumemto.mem`` of the reference's
  Python bindings (python_bindings/src/mumemto_pybind.cpp:93-118 over mumemto_library/mumemto_api.cpp:332-411).
* ``find_files`` / ``write_outputs``: what ``mumemto_exec`` does for FASTA paths (src/pfp_mum.cpp:31-159), writing
  PREFIX.mums | .mems | .bumbl, PREFIX.lengths, PREFIX.athresh, PREFIX.sa/.lcp/.bwt byte-for-byte like the reference.
"""
import os

import numpy as np

from ._binding import Engine


class MumResult:
    """lengths[n] uint32, offsets[n][num_docs] int64 (-1 absent), strands[n][num_docs] uint8 (1 = '+'), record lengths per doc."""

    def __init__(self, lengths, offsets, strands, record_lengths):
        self.lengths, self.offsets, self.strands, self.record_lengths = lengths, offsets, strands, record_lengths

    def __len__(self):
        return len(self.lengths)

    def match_at(self, i):
        return int(self.lengths[i]), self.offsets[i], self.strands[i]


class MemResult:
    def __init__(self, lengths, first, offsets, seq_ids, strands, record_lengths):
        self.lengths, self.first, self.offsets, self.seq_ids, self.strands, self.record_lengths = lengths, first, offsets, seq_ids, strands, record_lengths

    def __len__(self):
        return len(self.lengths)

    def match_at(self, i):
        a, b = int(self.first[i]), int(self.first[i + 1])
        return int(self.lengths[i]), self.offsets[a:b], self.seq_ids[a:b], self.strands[a:b]


def _record_lengths(sequences):
    return [[len(r) for r in doc] for doc in sequences]


def mum(sequences, min_match_len=20, use_revcomp=True, num_distinct=0, use_gsacak=False, device=0, _lib=None):
    """Multi-MUMs of in-memory sequences: sequences[doc][record] -> str.  use_gsacak is accepted for signature parity; both
    reference routes give the same result and the engine has a single device route."""
    if len(sequences) == 0:
        return MumResult(np.zeros(0, np.uint32), np.zeros((0, 0), np.int64), np.zeros((0, 0), np.uint8), [])
    e = Engine(device, _lib=_lib)
    try:
        e.load_records(sequences, use_revcomp)
        e.find(min_match_len=min_match_len, num_distinct_docs=num_distinct, rare_freq=1, max_mem_freq=0, cli_normalise=False, want_text=False)
        return MumResult(*e.mums(), _record_lengths(sequences))
    finally:
        e.close()


def mem(sequences, min_match_len=20, use_revcomp=True, num_distinct=0, max_total_freq=0, max_doc_freq=2, use_gsacak=False, device=0, _lib=None):
    if len(sequences) == 0:
        return MemResult(np.zeros(0, np.uint32), np.zeros(1, np.uint64), np.zeros(0, np.int64), np.zeros(0, np.uint64), np.zeros(0, np.uint8), [])
    if max_doc_freq <= 1:
        raise ValueError("per-sequence MEM frequency f must be > 1 (use mumemto_mum instead)")
    e = Engine(device, _lib=_lib)
    try:
        e.load_records(sequences, use_revcomp)
        e.find(min_match_len=min_match_len, num_distinct_docs=num_distinct, rare_freq=max_doc_freq, max_mem_freq=max_total_freq,
               cli_normalise=False, want_text=False)
        return MemResult(*e.mems(), _record_lengths(sequences))
    finally:
        e.close()


def _five_byte(a):
    """uint64 array -> 5 little-endian bytes per entry (THRBYTES/SSABYTES = 5, reference include/common.hpp:59-60)."""
    return np.ascontiguousarray(a.astype("<u8").view(np.uint8).reshape(-1, 8)[:, :5]).tobytes()


def write_outputs(engine, prefix, paths=None, binary=False, anchor_merge=False, arrays_out=False):
    """Write the reference's output files for the engine's last find()."""
    written = []

    def put(ext, data):
        with open(prefix + ext, "wb") as f:
            f.write(data)
        written.append(prefix + ext)

    if paths is not None:
        lines = []
        for d, p in enumerate(paths):
            cp = os.path.realpath(p)
            recs = engine.records(d)
            lines.append(f"{cp} * {sum(l for _, l in recs)}\n")
            lines += [f"{cp} {name} {l}\n" for name, l in recs]
        put(".lengths", "".join(lines).encode())
    if not engine.mum_mode:
        put(".mems", engine.text())
    elif binary:
        if engine.num_matches == 0:
            raise RuntimeError("zero MUMs with binary output: the reference dereferences an empty vector here (include/mem_finder.hpp:461-462)")
        partial = engine.stats()  # flags need num_distinct < num_docs; carried by the caller through find()
        put(".bumbl", engine.bumbl_bytes(getattr(engine, "_partial", False)))
    else:
        put(".mums", engine.text())
    if anchor_merge:
        t = engine.thresholds()
        put(".athresh", t[: len(t) // 2].tobytes())
    if arrays_out:
        sa, lcp, bwt = engine.arrays()
        put(".sa", _five_byte(sa)); put(".lcp", _five_byte(lcp)); put(".bwt", bwt.tobytes())
    return written


def find_files(paths, prefix, min_match_len=20, num_distinct_docs=0, rare_freq=1, max_mem_freq=0, revcomp=True, binary=False, merge=False,
               anchor_merge=False, arrays_out=False, device=0, route="auto", tile_ranks=0, no_prune=False, trim_min_len=0, shard=None, _lib=None):
    """`mumemto_exec [flags] -o prefix paths...` through the Python binding (the C++ CLI in csrc/cli_main.cpp is the drop-in).
    shard = (index, count): this call writes only the slice `index` of the match list; `cat` of the slices = the whole."""
    uniq, seen = [], set()
    for p in paths:
        ap = os.path.normpath(os.path.abspath(p))
        if ap not in seen:
            seen.add(ap); uniq.append(ap)
    e = Engine(device, _lib=_lib)
    try:
        e.load_fasta_files(uniq, revcomp)
        if binary and rare_freq != 1:
            binary = False
        st = e.find(min_match_len=min_match_len, num_distinct_docs=num_distinct_docs, rare_freq=rare_freq, max_mem_freq=max_mem_freq,
                    merge=merge, anchor_merge=anchor_merge, want_text=not binary, route=route, keep_arrays=arrays_out, tile_ranks=tile_ranks, no_prune=no_prune, trim_min_len=trim_min_len,
                    shard=shard)
        N = e.num_docs
        k = num_distinct_docs
        k = 2 if k < -N else (N + k if k <= 0 else (2 if k == 1 else min(k, N)))
        e._partial = k < N
        write_outputs(e, prefix, uniq, binary=binary, anchor_merge=anchor_merge, arrays_out=arrays_out)
        return st
    finally:
        e.close()


def read_lengths(path, revcomp=True):
    """PREFIX.lengths -> forward bases per document, the way RefBuilder(prefix) reads it (src/ref_builder.cpp:140-168): lines
    `path * total` (or the two-word form `path total`) count, per-record lines are skipped."""
    out = []
    with open(path) as f:
        for line in f:
            w = line.rstrip("\n").split(" ")
            if len(w) == 2:
                out.append(int(w[1]))
            elif len(w) == 3 and w[1] == "*":
                out.append(int(w[2]))
    return out


def find_arrays(arrays_prefix, prefix, min_match_len=20, num_distinct_docs=0, rare_freq=1, max_mem_freq=0, revcomp=True, binary=False, merge=False,
                anchor_merge=False, device=0, tile_ranks=0, _lib=None):
    """`mumemto_exec -a arrays_prefix [flags] -o prefix`: detection over arrays_prefix.{lengths,sa,lcp,bwt} (reference src/pfp_mum.cpp:96-110,
    include/read_arrays.hpp:66-84 -- which stops after |T| of the |T|+1 entries)."""
    lengths = read_lengths(arrays_prefix + ".lengths", revcomp)
    sa5 = np.fromfile(arrays_prefix + ".sa", dtype=np.uint8); lcp5 = np.fromfile(arrays_prefix + ".lcp", dtype=np.uint8)
    bwt = np.fromfile(arrays_prefix + ".bwt", dtype=np.uint8)
    total = sum((2 if revcomp else 1) * (l + 1) for l in lengths)
    if len(sa5) < 5 * total or len(lcp5) < 5 * total or len(bwt) < total:
        raise ValueError("array files are shorter than the sequence lengths imply")
    e = Engine(device, _lib=_lib)
    try:
        if binary and rare_freq != 1:
            binary = False
        st = e.find_from_arrays(lengths, sa5, lcp5, bwt, total, revcomp=revcomp, min_match_len=min_match_len, num_distinct_docs=num_distinct_docs,
                                rare_freq=rare_freq, max_mem_freq=max_mem_freq, merge=merge, anchor_merge=anchor_merge, want_text=not binary,
                                tile_ranks=tile_ranks)
        N = e.num_docs
        k = num_distinct_docs
        k = 2 if k < -N else (N + k if k <= 0 else (2 if k == 1 else min(k, N)))
        e._partial = k < N
        write_outputs(e, prefix, None, binary=binary, anchor_merge=anchor_merge)
        return st
    finally:
        e.close()


def read_mums(path):
    """Strict .mums text -> (lengths u32[n], offsets i64[n][N], strands u8[n][N] with 1 = '+'); include/mumsio.hpp:40-103."""
    lengths, offsets, strands = [], [], []
    with open(path) as f:
        for line in f:
            if not line.strip():
                continue
            a, b, c = line.split()[:3]
            o = b.split(",")
            if any(x == "" for x in o):
                raise ValueError("Cannot parse partial MUMs: empty offset encountered. Filter to strict MUMs.")
            lengths.append(int(a)); offsets.append([int(x) for x in o]); strands.append([1 if x == "+" else 0 for x in c.split(",")])
    N = len(offsets[0]) if offsets else 2
    return (np.array(lengths, dtype=np.uint32), np.array(offsets, dtype=np.int64).reshape(-1, N), np.array(strands, dtype=np.uint8).reshape(-1, N))


def merge_files(mums_paths, out_prefix, device=0, _lib=None):
    """`anchor_merge p1.mums p2.mums ... -o out_prefix` (src/merge_candidates.cpp:170-255) through the Python binding."""
    parts = []
    for p in mums_paths:
        l, o, s = read_mums(p)
        t = np.fromfile(p[:-5] + ".athresh", dtype=np.uint16)
        parts.append((l, o, s, t, len(l), o.shape[1]))
    anchor_len = len(parts[0][3])
    if any(len(p[3]) != anchor_len for p in parts):
        raise ValueError("threshold files differ in length: the partitions do not share their first sequence")
    e = Engine(device, _lib=_lib)
    try:
        e.anchor_merge(parts, anchor_len)
        with open(out_prefix + ".mums", "wb") as f:
            f.write(e.merge_text())
        e.merge_result()[3].tofile(out_prefix + ".athresh")
    finally:
        e.close()
