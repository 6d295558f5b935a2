"""ctypes binding of include/mumemto_b200.h (libmumemto_b200.so, CUDA sm_100a).

There is no CPU implementation behind this module: if the shared library has not been built, or no CUDA device is present,
construction of an Engine raises.  (tests/ can inject the SIMT unit-test build of the same kernel sources through the private
``_lib`` argument; nothing in the package does.)
"""
import ctypes as C
import gzip
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmumemto_b200.so")

u8p = C.POINTER(C.c_uint8)


class Options(C.Structure):
    _fields_ = [("min_match_len", C.c_uint32), ("num_distinct_docs", C.c_int32), ("rare_freq", C.c_int32), ("max_mem_freq", C.c_int32),
                ("merge", C.c_uint8), ("anchor_merge", C.c_uint8), ("cli_normalise", C.c_uint8), ("want_text", C.c_uint8),
                ("route", C.c_uint8), ("keep_arrays", C.c_uint8), ("no_prune", C.c_uint8), ("shard_count", C.c_uint8), ("shard_index", C.c_uint8), ("reserved", C.c_uint8 * 3), ("tile_ranks", C.c_uint64), ("trim_min_len", C.c_uint32), ("reserved2", C.c_uint32)]


class Stats(C.Structure):
    _fields_ = [("text_symbols", C.c_uint64), ("input_bases", C.c_uint64), ("num_matches", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("sa_rounds", C.c_uint32), ("sa_sorted_elements", C.c_uint64),
                ("ms_text", C.c_float), ("ms_sa", C.c_float), ("ms_lcp", C.c_float), ("ms_detect", C.c_float), ("ms_emit", C.c_float),
                ("ms_total", C.c_float), ("route", C.c_uint32), ("pfp_phrases", C.c_uint32), ("pfp_distinct_phrases", C.c_uint32),
                ("pfp_dict_bytes", C.c_uint32), ("pfp_tiles", C.c_uint32), ("pruned", C.c_uint32), ("stream_suffixes", C.c_uint64),
                ("threshold_segments", C.c_uint64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class MumemtoError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[{code}] {msg}")
        self.code = code


_SIGNATURES = {
    "mmb_last_error": (C.c_char_p, []),
    "mmb_default_options": (None, [C.POINTER(Options)]),
    "mmb_create": (C.c_int, [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "mmb_destroy": (None, [C.c_void_p]),
    "mmb_alloc_pinned": (C.c_void_p, [C.c_size_t]),
    "mmb_free_pinned": (None, [C.c_void_p]),
    "mmb_load_fasta": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.c_size_t, C.c_int]),
    "mmb_load_fasta_device": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_size_t, C.c_int]),
    "mmb_load_records": (C.c_int, [C.c_void_p, C.POINTER(C.c_char_p), C.POINTER(C.c_size_t), C.c_size_t, C.c_int]),
    "mmb_find": (C.c_int, [C.c_void_p, C.POINTER(Options)]),
    "mmb_get_stats": (C.c_int, [C.c_void_p, C.POINTER(Stats)]),
    "mmb_num_docs": (C.c_size_t, [C.c_void_p]),
    "mmb_doc_lengths": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "mmb_num_records": (C.c_size_t, [C.c_void_p, C.c_size_t]),
    "mmb_record": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t, C.POINTER(C.c_char_p), C.POINTER(C.c_uint64)]),
    "mmb_is_mum_mode": (C.c_int, [C.c_void_p]),
    "mmb_num_matches": (C.c_uint64, [C.c_void_p]),
    "mmb_text_size": (C.c_uint64, [C.c_void_p]),
    "mmb_copy_text": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mmb_text_device": (C.c_void_p, [C.c_void_p]),
    "mmb_copy_mums": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mmb_copy_mum_strand_bits": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mmb_num_mem_occurrences": (C.c_uint64, [C.c_void_p]),
    "mmb_copy_mems": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mmb_threshold_count": (C.c_uint64, [C.c_void_p]),
    "mmb_copy_thresholds": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mmb_copy_arrays": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mmb_copy_symbols": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mmb_profile_kernel": (C.c_int, [C.c_char_p]),
    "mmb_find_from_arrays": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                                       C.POINTER(Options)]),
    "mmb_profile_result": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
    "mmb_profile_table": (C.c_int, [C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "mmb_mum_table_device": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "mmb_anchor_merge": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.POINTER(C.c_void_p),
                                   C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.c_uint64, C.c_int]),
    "mmb_merge_num_mums": (C.c_uint64, [C.c_void_p]),
    "mmb_merge_num_docs": (C.c_uint64, [C.c_void_p]),
    "mmb_merge_copy": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mmb_merge_text_size": (C.c_uint64, [C.c_void_p]),
    "mmb_merge_copy_text": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mmb_merge_text_device": (C.c_void_p, [C.c_void_p]),
}


def declare(lib):
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    return lib


_lib_cache = None


def load_library():
    """The CUDA library, or a loud failure."""
    global _lib_cache
    if _lib_cache is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is not built: run `make` (nvcc, sm_100a). mumemto_b200 has no CPU fallback.")
        _lib_cache = declare(C.CDLL(LIB_PATH))
    return _lib_cache


def read_fasta_image(path):
    """File bytes as the reference sees them through gzopen/gzread (plain or .gz)."""
    with open(path, "rb") as f:
        head = f.read(2)
    if head == b"\x1f\x8b":
        with gzip.open(path, "rb") as f:
            return f.read()
    with open(path, "rb") as f:
        return f.read()


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Engine:
    def __init__(self, device=0, stream=None, _lib=None):
        self.lib = _lib if _lib is not None else load_library()
        h = C.c_void_p()
        self._check(self.lib.mmb_create(device, C.c_void_p(stream) if stream else None, C.byref(h)))
        self.h = h
        self._keep = None

    def _check(self, rc):
        if rc != 0:
            raise MumemtoError(rc, self.lib.mmb_last_error().decode(errors="replace"))

    def close(self):
        if getattr(self, "h", None):
            self.lib.mmb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ input
    def load_fasta_bytes(self, images, revcomp=True):
        """images: list of bytes-like / uint8 numpy arrays, one decompressed FASTA file each (host memory)."""
        arrs = [np.frombuffer(b, dtype=np.uint8) if not isinstance(b, np.ndarray) else b for b in images]
        n = len(arrs)
        ptrs = (C.c_void_p * n)(*[a.ctypes.data if a.size else None for a in arrs])
        sizes = (C.c_uint64 * n)(*[a.size for a in arrs])
        self._keep = arrs
        self._check(self.lib.mmb_load_fasta(self.h, ptrs, sizes, n, int(revcomp)))

    def load_fasta_files(self, paths, revcomp=True):
        self.load_fasta_bytes([read_fasta_image(p) for p in paths], revcomp)

    def load_fasta_device(self, dev_ptr, offsets, sizes, revcomp=True):
        n = len(offsets)
        self._check(self.lib.mmb_load_fasta_device(self.h, C.c_void_p(dev_ptr), (C.c_uint64 * n)(*offsets), (C.c_uint64 * n)(*sizes), n, int(revcomp)))

    def load_records(self, docs, revcomp=True):
        flat = [r.encode() if isinstance(r, str) else bytes(r) for d in docs for r in d]
        recs = (C.c_char_p * max(1, len(flat)))(*flat)
        counts = (C.c_size_t * max(1, len(docs)))(*[len(d) for d in docs])
        self._check(self.lib.mmb_load_records(self.h, recs, counts, len(docs), int(revcomp)))

    # ------------------------------------------------------------------ hot path
    ROUTES = {"auto": 0, "direct": 1, "pfp": 2}

    def find(self, min_match_len=20, num_distinct_docs=0, rare_freq=1, max_mem_freq=0, merge=False, anchor_merge=False, cli_normalise=True,
             want_text=True, route="auto", keep_arrays=False, tile_ranks=0, no_prune=False, trim_min_len=0, shard=None):
        """shard = (index, count): only the matches of that slice of the suffix order; outputs concatenate in index order."""
        o = Options(min_match_len, num_distinct_docs, rare_freq, max_mem_freq, int(merge), int(anchor_merge), int(cli_normalise), int(want_text),
                    self.ROUTES[route], int(keep_arrays), int(no_prune), shard[1] if shard else 0, shard[0] if shard else 0, (C.c_uint8 * 3)(),
                    int(tile_ranks), int(trim_min_len), 0)
        self._check(self.lib.mmb_find(self.h, C.byref(o)))
        return self.stats()

    def find_from_arrays(self, doc_lengths, sa5, lcp5, bwt, entries, revcomp=True, min_match_len=20, num_distinct_docs=0, rare_freq=1,
                         max_mem_freq=0, merge=False, anchor_merge=False, cli_normalise=True, want_text=True, tile_ranks=0):
        """Detection over pre-computed arrays in the reference's file format (uint8 arrays: 5-byte LE SA / LCP entries, BWT bytes)."""
        o = Options(min_match_len, num_distinct_docs, rare_freq, max_mem_freq, int(merge), int(anchor_merge), int(cli_normalise), int(want_text),
                    0, 0, 0, 0, 0, (C.c_uint8 * 3)(), int(tile_ranks), 0, 0)
        dl = np.ascontiguousarray(doc_lengths, dtype=np.uint64)
        self._keep_arrays = (dl, sa5, lcp5, bwt)
        self._check(self.lib.mmb_find_from_arrays(self.h, dl.ctypes.data_as(C.POINTER(C.c_uint64)), len(dl), int(revcomp), _ptr(sa5), _ptr(lcp5),
                                                  _ptr(bwt), int(entries), C.byref(o)))
        return self.stats()

    def stats(self):
        s = Stats()
        self._check(self.lib.mmb_get_stats(self.h, C.byref(s)))
        return s.as_dict()

    # ------------------------------------------------------------------ results
    @property
    def num_docs(self):
        return self.lib.mmb_num_docs(self.h)

    def doc_lengths(self):
        a = np.zeros(self.num_docs, dtype=np.uint64)
        self._check(self.lib.mmb_doc_lengths(self.h, a.ctypes.data_as(C.POINTER(C.c_uint64))))
        return a

    def records(self, doc):
        out = []
        for r in range(self.lib.mmb_num_records(self.h, doc)):
            name, length = C.c_char_p(), C.c_uint64()
            self._check(self.lib.mmb_record(self.h, doc, r, C.byref(name), C.byref(length)))
            out.append(((name.value or b"").decode(errors="replace"), length.value))
        return out

    @property
    def mum_mode(self):
        return bool(self.lib.mmb_is_mum_mode(self.h))

    @property
    def num_matches(self):
        return self.lib.mmb_num_matches(self.h)

    def text(self):
        n = self.lib.mmb_text_size(self.h)
        buf = np.empty(n, dtype=np.uint8)
        self._check(self.lib.mmb_copy_text(self.h, _ptr(buf)))
        return buf.tobytes()

    def copy_text_into(self, host_array):
        self._check(self.lib.mmb_copy_text(self.h, _ptr(host_array)))
        return self.lib.mmb_text_size(self.h)

    def mums(self):
        n, N = self.num_matches, self.num_docs
        lengths = np.empty(n, dtype=np.uint32); offsets = np.empty((n, N), dtype=np.int64); strands = np.empty((n, N), dtype=np.uint8)
        self._check(self.lib.mmb_copy_mums(self.h, _ptr(lengths), _ptr(offsets), _ptr(strands)))
        return lengths, offsets, strands

    def mum_strand_bits(self):
        bits = np.empty((self.num_matches * self.num_docs + 7) // 8, dtype=np.uint8)
        self._check(self.lib.mmb_copy_mum_strand_bits(self.h, _ptr(bits)))
        return bits

    def mems(self):
        n, occ = self.num_matches, self.lib.mmb_num_mem_occurrences(self.h)
        lengths = np.empty(n, dtype=np.uint32); first = np.empty(n + 1, dtype=np.uint64)
        offsets = np.empty(occ, dtype=np.int64); seq_ids = np.empty(occ, dtype=np.uint64); strands = np.empty(occ, dtype=np.uint8)
        self._check(self.lib.mmb_copy_mems(self.h, _ptr(lengths), _ptr(first), _ptr(offsets), _ptr(seq_ids), _ptr(strands)))
        return lengths, first, offsets, seq_ids, strands

    def thresholds(self):
        t = np.empty(self.lib.mmb_threshold_count(self.h), dtype=np.uint16)
        self._check(self.lib.mmb_copy_thresholds(self.h, _ptr(t)))
        return t

    def arrays(self):
        m = int(self.stats()["text_symbols"]) + 1
        sa = np.empty(m, dtype=np.uint64); lcp = np.empty(m, dtype=np.uint64); bwt = np.empty(m, dtype=np.uint8)
        self._check(self.lib.mmb_copy_arrays(self.h, _ptr(sa), _ptr(lcp), _ptr(bwt)))
        return sa, lcp, bwt

    def symbols(self):
        t = np.empty(int(self.stats()["text_symbols"]), dtype=np.uint8)
        self._check(self.lib.mmb_copy_symbols(self.h, _ptr(t)))
        return t.tobytes()

    def bumbl_bytes(self, partial):
        """The .bumbl image the reference assembles in mem_finder::write_bums (include/mem_finder.hpp:451-503)."""
        lengths, offsets, _ = self.mums()
        flags = np.uint16(0x8000 | (0x2000 if partial else 0))
        head = flags.tobytes() + np.uint64(self.num_docs).tobytes() + np.uint64(len(lengths)).tobytes()
        return head + lengths.tobytes() + offsets.tobytes() + self.mum_strand_bits().tobytes()

    # ------------------------------------------------------------------ measurement aid
    def profile_kernel(self, name_substring):
        self.lib.mmb_profile_kernel(name_substring.encode() if name_substring else None)

    def profile_result(self):
        ms, n = C.c_double(), C.c_uint64()
        self.lib.mmb_profile_result(C.byref(ms), C.byref(n))
        return ms.value, n.value

    def profile_table(self):
        """{kernel name: (ms, launches)} of everything recorded since profile_kernel("*")."""
        need = C.c_size_t()
        self.lib.mmb_profile_table(None, 0, C.byref(need))
        buf = C.create_string_buffer(need.value + 16)
        self.lib.mmb_profile_table(buf, need.value + 16, C.byref(need))
        out = {}
        for line in buf.value.decode().splitlines():
            name, ms, n = line.split("\t")
            out[name] = (float(ms), int(n))
        return out

    # ------------------------------------------------------------------ partition merge
    def mum_table_device(self):
        """Device pointers (lengths, offsets, strands, thresholds) of the last MUM table, as ints."""
        p = [C.c_void_p() for _ in range(4)]
        self._check(self.lib.mmb_mum_table_device(self.h, *[C.byref(x) for x in p]))
        return tuple(x.value for x in p)

    def anchor_merge(self, parts, anchor_len, on_device=False):
        """parts: per partition (lengths, offsets, strands, thresholds, n_mums, ndocs); host numpy arrays, or device pointers (ints)
        when on_device.  Folded left to right like `anchor_merge p1 p2 ...`."""
        k = len(parts)
        def col(i):
            return (C.c_void_p * k)(*[(p[i] if on_device else p[i].ctypes.data) for p in parts])
        n = (C.c_uint64 * k)(*[int(p[4]) for p in parts])
        nd = (C.c_uint64 * k)(*[int(p[5]) for p in parts])
        self._keep = parts
        self._check(self.lib.mmb_anchor_merge(self.h, k, col(0), n, col(1), col(2), col(3), nd, int(anchor_len), int(on_device)))
        self._merge_anchor_len = int(anchor_len)

    def merge_result(self):
        n, nd = self.lib.mmb_merge_num_mums(self.h), self.lib.mmb_merge_num_docs(self.h)
        lengths = np.empty(n, dtype=np.uint32); offsets = np.empty((n, nd), dtype=np.int64); strands = np.empty((n, nd), dtype=np.uint8)
        thresholds = np.empty(self._merge_anchor_len, dtype=np.uint16)
        self._check(self.lib.mmb_merge_copy(self.h, _ptr(lengths), _ptr(offsets), _ptr(strands), _ptr(thresholds)))
        return lengths, offsets, strands, thresholds

    def merge_text(self):
        buf = np.empty(self.lib.mmb_merge_text_size(self.h), dtype=np.uint8)
        self._check(self.lib.mmb_merge_copy_text(self.h, _ptr(buf)))
        return buf.tobytes()

    def merge_text_size(self):
        return int(self.lib.mmb_merge_text_size(self.h))

    def merge_text_into(self, host_array):
        """Merged .mums bytes into a caller-owned (ideally pinned) uint8 array of at least merge_text_size() bytes; returns the size."""
        n = self.merge_text_size()
        if host_array.nbytes < n:
            raise ValueError("destination too small for the merged text")
        self._check(self.lib.mmb_merge_copy_text(self.h, _ptr(host_array)))
        return n
