"""Deterministic synthetic phage-display-shaped peptide sets.

F = N/40 motif families; each family = a random parent over the 20 standard residues; 60 % of
the sequences are family members (parent with 0-4 substitutions and, with probability 0.25, a
shift by s in {-2,-1,+1,+2}), 40 % uniform-random background; duplicates rejected; abundance
Zipf(s = 1.1) by rank with the family parents ranked first.  Output is ALREADY in Hammock's
clustering order ("size": abundance desc, then string desc; UniqueSequence.java:176-181).
numpy PCG64 seeded explicitly, so every process regenerates identical bytes.
"""
from __future__ import annotations

import numpy as np

ALPHABET = "ARNDCQEGHILKMFPSTWYVBZX*"
_ASCII = np.frombuffer(ALPHABET.encode(), dtype=np.uint8)
BLOSUM62_ROWS = """
 4 -1 -2 -2  0 -1 -1  0 -2 -1 -1 -1 -1 -2 -1  1  0 -3 -2  0 -2 -1  0 -4
-1  5  0 -2 -3  1  0 -2  0 -3 -2  2 -1 -3 -2 -1 -1 -3 -2 -3 -1  0 -1 -4
-2  0  6  1 -3  0  0  0  1 -3 -3  0 -2 -3 -2  1  0 -4 -2 -3  3  0 -1 -4
-2 -2  1  6 -3  0  2 -1 -1 -3 -4 -1 -3 -3 -1  0 -1 -4 -3 -3  4  1 -1 -4
 0 -3 -3 -3  9 -3 -4 -3 -3 -1 -1 -3 -1 -2 -3 -1 -1 -2 -2 -1 -3 -3 -2 -4
-1  1  0  0 -3  5  2 -2  0 -3 -2  1  0 -3 -1  0 -1 -2 -1 -2  0  3 -1 -4
-1  0  0  2 -4  2  5 -2  0 -3 -3  1 -2 -3 -1  0 -1 -3 -2 -2  1  4 -1 -4
 0 -2  0 -1 -3 -2 -2  6 -2 -4 -4 -2 -3 -3 -2  0 -2 -2 -3 -3 -1 -2 -1 -4
-2  0  1 -1 -3  0  0 -2  8 -3 -3 -1 -2 -1 -2 -1 -2 -2  2 -3  0  0 -1 -4
-1 -3 -3 -3 -1 -3 -3 -4 -3  4  2 -3  1  0 -3 -2 -1 -3 -1  3 -3 -3 -1 -4
-1 -2 -3 -4 -1 -2 -3 -4 -3  2  4 -2  2  0 -3 -2 -1 -2 -1  1 -4 -3 -1 -4
-1  2  0 -1 -3  1  1 -2 -1 -3 -2  5 -1 -3 -1  0 -1 -3 -2 -2  0  1 -1 -4
-1 -1 -2 -3 -1  0 -2 -3 -2  1  2 -1  5  0 -2 -1 -1 -1 -1  1 -3 -1 -1 -4
-2 -3 -3 -3 -2 -3 -3 -3 -1  0  0 -3  0  6 -4 -2 -2  1  3 -1 -3 -3 -1 -4
-1 -2 -2 -1 -3 -1 -1 -2 -2 -3 -3 -1 -2 -4  7 -1 -1 -4 -3 -2 -2 -1 -2 -4
 1 -1  1  0 -1  0  0  0 -1 -2 -2  0 -1 -2 -1  4  1 -3 -2 -2  0  0  0 -4
 0 -1  0 -1 -1 -1 -1 -2 -2 -1 -1 -1 -1 -2 -1  1  5 -2 -2  0 -1 -1  0 -4
-3 -3 -4 -4 -2 -2 -3 -2 -2 -3 -2 -3 -1  1 -4 -3 -2 11  2 -3 -4 -3 -2 -4
-2 -2 -2 -3 -2 -1 -2 -3  2 -1 -1 -2 -1  3 -3 -2 -2  2  7 -1 -3 -2 -1 -4
 0 -3 -3 -3 -1 -2 -2 -3 -3  3  1 -2  1 -1 -2 -2  0 -3 -1  4 -3 -2 -1 -4
-2 -1  3  4 -3  0  1 -1  0 -3 -4  0 -3 -3 -2  0 -1 -4 -3 -3  4  1 -1 -4
-1  0  0  1 -3  3  4 -2  0 -3 -3  1 -1 -3 -1  0 -1 -3 -2 -2  1  4 -1 -4
 0 -1 -1 -1 -2 -1 -1 -1 -1 -1 -1 -1 -1 -1 -2  0  0 -2 -1 -1 -1 -1 -1 -4
-4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4  1
"""


def blosum62() -> np.ndarray:
    """The standard BLOSUM62 table in Hammock's residue order (public NCBI matrix; equals what
    FileIOManager.loadScoringMatrix returns for matrices/blosum62.txt, checked in tests)."""
    return np.array(BLOSUM62_ROWS.split(), dtype=np.int32).reshape(24, 24)


def java_round(x: float) -> int:
    return int(np.floor(x + 0.5))


def default_params(lengths: np.ndarray):
    """Hammock.java:394-401, 1409-1434: T = round(1.7 mean), X = min(round(mean/4), minLen-1), K = round(0.025 N)"""
    mean = float(lengths.sum()) / float(len(lengths))
    return java_round(mean * 1.7), min(java_round(mean / 4), int(lengths.min()) - 1), java_round(len(lengths) * 0.025)


def generate(n: int, min_len: int = 12, max_len: int = 12, seed: int = 20260101, zipf_s: float = 1.1,
             top_abundance: int = 100000):
    """-> dict(residues u8, offsets i32[n+1], abundance i32[n], lengths) in clustering order."""
    rng = np.random.Generator(np.random.PCG64(seed))
    nfam = max(1, n // 40)
    Lmax = max_len
    PAD = 255

    def random_rows(m, lens):
        rows = rng.integers(0, 20, size=(m, Lmax), dtype=np.uint8)
        rows[np.arange(Lmax)[None, :] >= lens[:, None]] = PAD
        return rows

    fam_len = rng.integers(min_len, max_len + 1, size=nfam)
    parents = random_rows(nfam, fam_len)

    def members(m):
        f = rng.integers(0, nfam, size=m)
        rows = parents[f].copy()
        lens = fam_len[f]
        nsub = rng.integers(0, 5, size=m)
        for s in range(4):
            sel = np.nonzero(nsub > s)[0]
            pos = (rng.random(len(sel)) * lens[sel]).astype(np.int64)
            rows[sel, pos] = rng.integers(0, 20, size=len(sel), dtype=np.uint8)
        shift = rng.random(m) < 0.25
        sh = rng.choice(np.array([-2, -1, 1, 2]), size=m)
        for s in (-2, -1, 1, 2):
            sel = np.nonzero(shift & (sh == s))[0]
            if len(sel) == 0:
                continue
            a = abs(s)
            for L in range(min_len, max_len + 1):
                sl = sel[lens[sel] == L]
                if len(sl) == 0:
                    continue
                blk = rows[sl, :L]
                fresh = rng.integers(0, 20, size=(len(sl), a), dtype=np.uint8)
                if s > 0:    # drop a residues at the left end, append a random ones at the right
                    blk = np.concatenate([blk[:, a:], fresh], axis=1)
                else:
                    blk = np.concatenate([fresh, blk[:, :L - a]], axis=1)
                rows[sl, :L] = blk
        return rows, lens

    def background(m):
        lens = rng.integers(min_len, max_len + 1, size=m)
        return random_rows(m, lens), lens

    def keys(rows):   # injective key per (padded) row, for duplicate rejection
        k = np.zeros(len(rows), dtype=object) if Lmax > 12 else None
        if k is None:
            k = np.zeros(len(rows), dtype=np.uint64)
            for j in range(Lmax):
                k = k * np.uint64(32) + (rows[:, j].astype(np.uint64) & np.uint64(31))
            return k
        return np.array([r.tobytes() for r in rows], dtype=object)

    rows = parents.copy()
    lens = fam_len.copy()
    kind = np.zeros(nfam, dtype=np.uint8)          # 0 parent, 1 member, 2 background
    seen = {}
    # de-duplicate parents first
    k = keys(rows)
    _, first = np.unique(k, return_index=True)
    first.sort()
    rows, lens, kind = rows[first], lens[first], kind[first]
    seen_keys = set(keys(rows).tolist())
    n_mem_target = int(0.6 * n)
    while len(rows) < n:
        need = n - len(rows)
        n_mem_have = int((kind <= 1).sum())
        m_mem = min(need, max(0, n_mem_target - n_mem_have))
        m_bg = need - m_mem
        parts = []
        if m_mem:
            r, l = members(int(m_mem * 1.1) + 16)
            parts.append((r, l, 1, m_mem))
        if m_bg:
            r, l = background(int(m_bg * 1.02) + 16)
            parts.append((r, l, 2, m_bg))
        for r, l, kd, want in parts:
            kk = keys(r)
            _, fi = np.unique(kk, return_index=True)
            fi.sort()
            take = []
            for i in fi:
                key = kk[i].item() if hasattr(kk[i], "item") else kk[i]
                if key in seen_keys:
                    continue
                seen_keys.add(key)
                take.append(i)
                if len(take) == want:
                    break
            take = np.array(take, dtype=np.int64)
            rows = np.concatenate([rows, r[take]])
            lens = np.concatenate([lens, l[take]])
            kind = np.concatenate([kind, np.full(len(take), kd, dtype=np.uint8)])
    rows, lens, kind = rows[:n], lens[:n], kind[:n]
    # ranks: parents first, everything else shuffled
    rest = np.nonzero(kind != 0)[0]
    rng.shuffle(rest)
    rank_order = np.concatenate([np.nonzero(kind == 0)[0], rest])
    ab = np.maximum(1, np.floor(top_abundance / np.power(np.arange(1, n + 1, dtype=np.float64), zipf_s))).astype(np.int64)
    abundance = np.empty(n, dtype=np.int32)
    abundance[rank_order] = ab.astype(np.int32)
    # clustering order: abundance desc, then upper-case string desc (shorter prefix sorts lower)
    letters = np.where(rows == PAD, 0, _ASCII[np.minimum(rows, 23)]).astype(np.uint8)
    sort_keys = [letters[:, j] for j in range(Lmax - 1, -1, -1)] + [abundance]
    asc = np.lexsort(sort_keys)
    order = asc[::-1]
    rows, lens, abundance = rows[order], lens[order].astype(np.int32), abundance[order]
    offsets = np.zeros(n + 1, dtype=np.int32)
    offsets[1:] = np.cumsum(lens)
    residues = rows[rows != PAD] if min_len != max_len else rows.reshape(-1)
    residues = np.ascontiguousarray(rows[np.arange(Lmax)[None, :] < lens[:, None]], dtype=np.uint8)
    return {"residues": residues, "offsets": offsets, "abundance": np.ascontiguousarray(abundance),
            "lengths": lens}


def to_strings(residues, offsets):
    a = _ASCII[residues].tobytes().decode()
    return [a[offsets[i]:offsets[i + 1]] for i in range(len(offsets) - 1)]
