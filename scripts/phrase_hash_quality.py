#!/usr/bin/env python
"""Collision check of the phrase content hash (csrc/pfp.cu: phrase_hash) on pangenome-like phrases, emulated with Python integers.

The engine groups phrases by a 64-bit content hash and verifies every group byte by byte (a collision costs a retry with another seed,
never a wrong result) -- but a hash with a structural weakness collides on EVERY seed once there are millions of distinct phrases
(seen at 64 x 100 Mbp with a mix whose multiplications only carried differences upward: "phrase hashing kept colliding").  This script
cuts the bench generator's haplotypes into content-defined phrases, dedupes them exactly, and counts distinct phrases that share a hash,
for the mixes listed in MIXES (`words` has the structure of the kernel's hash: per-word finalisers, summed).  Usage: python scripts/phrase_hash_quality.py [N] [L]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

M = (1 << 64) - 1
PHI, K1, K2 = 0x9E3779B97F4A7C15, 0xbf58476d1ce4e5b9, 0x94d049bb133111eb


def mix64(x):
    x ^= x >> 30; x = (x * K1) & M; x ^= x >> 27; x = (x * K2) & M; x ^= x >> 31
    return x


def mix_weak(lo, hi, salt):                 # the mix that failed at full size
    u = ((lo ^ salt) * K1) & M; u ^= u >> 32
    v = ((hi + u) * K2) & M; v ^= v >> 29
    return (v * PHI) & M


def mix_mum(lo, hi, salt):                  # 64 x 64 -> 128 multiply, halves folded (a sound cheap alternative; not faster on the GPU)
    a = lo ^ salt
    b = hi ^ (((salt << 32) | (salt >> 32)) & M) ^ K2
    p = a * b
    return (p & M) ^ (p >> 64)


def mix_words(lo, hi, salt):                # what pfp.cu uses: a full 64-bit finaliser per 8-byte word (salted per word)
    return (mix64((lo + salt) & M) + mix64((hi + salt + (PHI >> 1)) & M)) & M


MIXES = {"weak": mix_weak, "mum": mix_mum, "words": mix_words}


def phrase_hash(p, seed, mix):
    h = seed ^ mix64(len(p))
    for c in range(0, len(p), 16):
        ch = p[c:c + 16].ljust(16, b"\0")
        lo, hi = int.from_bytes(ch[:8], "little"), int.from_bytes(ch[8:], "little")
        h = (h + mix(lo, hi, (seed + (c // 16 + 1) * PHI) & M)) & M
    return h


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
    imgs = bench.make_collection(dict(N=N, L=L, kind="pangenome"), seed=1)
    distinct = set()
    rng = np.random.default_rng(0)
    mult = rng.integers(1, 1 << 32, 10, dtype=np.uint64) | 1
    for im in imgs:
        raw = im.tobytes()
        seq = raw[raw.index(b"\n") + 1:].replace(b"\n", b"")
        a = np.frombuffer(seq, np.uint8).astype(np.uint64)
        hsh = np.zeros(len(a) - 9, np.uint64)
        for k in range(10):
            hsh += a[k:len(a) - 9 + k] * mult[k]
        hsh = (hsh * np.uint64(0x2C1B3C6D)) & np.uint64(0xffffffff)
        ends = np.flatnonzero(hsh < np.uint64(0xffffffff // 100)) + 9
        starts = np.concatenate(([0], ends - 9))
        stops = np.concatenate((ends + 1, [len(seq)]))
        for s, e in zip(starts, stops):
            distinct.add(seq[s:e])
    phrases = list(distinct)
    print(f"{N} x {L} bp: {len(phrases)} distinct phrases, mean length {sum(map(len, phrases)) / len(phrases):.1f}")
    for name, mix in MIXES.items():
        for seed in (mix64(0x1234567), mix64(0x1234568)):
            seen = {}
            coll = 0
            for p in phrases:
                h = phrase_hash(p, seed, mix)
                if h in seen:
                    coll += 1
                seen[h] = p
            print(f"  mix {name:5s} seed {seed:016x}: {coll} colliding phrases")


if __name__ == "__main__":
    main()
