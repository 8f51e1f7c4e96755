"""ctypes binding of the CPU oracle (oracle/hammock_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package hammock_b200 never imports it.
PARITY UNPINNED (see hammock_oracle.h): the reference is Java and cannot run here.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libhmkoracle.so")
ALPHABET = "ARNDCQEGHILKMFPSTWYVBZX*"

OK, ERR_SHIFT_TOO_BIG, ERR_NULL_CLUSTER, ERR_BAD_RESIDUE = 0, 1, 2, 3
ERR_FILE_FORMAT, ERR_IO = 6, 7
ERR_EMPTY, ERR_ASYMMETRIC, ERR_TREEIFIED = 8, 9, 10


class Counters(C.Structure):
    _fields_ = [
        ("p1_steps", C.c_int64), ("p1_new_clusters", C.c_int64), ("p1_joins", C.c_int64),
        ("p1_orphans", C.c_int64), ("p1_pairs", C.c_int64), ("p2_queries", C.c_int64),
        ("p2_assigned", C.c_int64), ("p2_pairs_early", C.c_int64), ("p2_pairs_dense", C.c_int64),
        ("cells", C.c_int64), ("npe_step", C.c_int32),
    ]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


class _Fasta(C.Structure):
    _fields_ = [("n", C.c_int32), ("residues", C.POINTER(C.c_uint8)), ("offsets", C.POINTER(C.c_int32)),
                ("abundance", C.POINTER(C.c_int32)), ("text", C.POINTER(C.c_char))]


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("hammock_oracle.c", "clinkage_oracle.c", "hammock_oracle.h")]
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        u8p, i32p = C.POINTER(C.c_uint8), C.POINTER(C.c_int32)
        L.hmko_score_with_shift.restype = C.c_int32
        L.hmko_score_with_shift.argtypes = [u8p, C.c_int, u8p, C.c_int, i32p, C.c_int, C.c_int,
                                            C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.hmko_pair_cells.restype = C.c_int64
        L.hmko_pair_cells.argtypes = [C.c_int, C.c_int, C.c_int]
        L.hmko_pair_shifts.restype = C.c_int
        L.hmko_pair_shifts.argtypes = [C.c_int, C.c_int, C.c_int]
        L.hmko_load_matrix.restype = C.c_int
        L.hmko_load_matrix.argtypes = [C.c_char_p, i32p, C.c_char_p, C.c_size_t]
        L.hmko_sort_order_size.restype = None
        L.hmko_sort_order_size.argtypes = [C.c_int32, u8p, i32p, i32p, i32p]
        L.hmko_default_params.restype = None
        L.hmko_default_params.argtypes = [C.c_int32, i32p, i32p, i32p, i32p]
        L.hmko_check_max_shift.restype = C.c_int32
        L.hmko_check_max_shift.argtypes = [C.c_int32, i32p, C.c_int32]
        L.hmko_greedy_cluster_bounded.restype = C.c_int
        L.hmko_greedy_cluster_bounded.argtypes = [C.c_int32, u8p, i32p, i32p, i32p, C.c_int32, C.c_int32,
                                                  C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int64,
                                                  i32p, i32p, i32p, i32p, i32p, C.POINTER(Counters)]
        L.hmko_clinkage_cluster.restype = C.c_int
        L.hmko_clinkage_cluster.argtypes = [C.c_int32, u8p, i32p, i32p, i32p, C.c_int32, C.c_int32, C.c_int32,
                                            i32p, i32p, i32p, i32p, C.POINTER(C.c_int64)]
        L.hmko_load_fasta.restype = C.c_int
        L.hmko_load_fasta.argtypes = [C.c_char_p, C.POINTER(_Fasta), C.c_char_p, C.c_size_t]
        L.hmko_fasta_free.restype = None
        L.hmko_fasta_free.argtypes = [C.POINTER(_Fasta)]
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def encode(s: str) -> np.ndarray:
    out = np.empty(len(s), dtype=np.uint8)
    for i, ch in enumerate(s.upper()):
        k = ALPHABET.find(ch)
        if k < 0:
            raise ValueError(f"character {ch} is not a valid letter from the amino acid alphabet code")
        out[i] = k
    return out


def pack(seqs):
    """list of str -> (residues u8, offsets i32[n+1])"""
    offs = np.zeros(len(seqs) + 1, dtype=np.int32)
    for i, s in enumerate(seqs):
        offs[i + 1] = offs[i] + len(s)
    res = np.empty(int(offs[-1]), dtype=np.uint8)
    for i, s in enumerate(seqs):
        res[offs[i]:offs[i + 1]] = encode(s)
    return res, offs


def load_matrix(path: str) -> np.ndarray:
    M = np.zeros(576, dtype=np.int32)
    err = C.create_string_buffer(256)
    rc = lib().hmko_load_matrix(path.encode(), _p(M, C.c_int32), err, 256)
    if rc:
        raise OracleError(rc, err.value.decode())
    return M.reshape(24, 24)


class OracleError(Exception):
    def __init__(self, status, msg=""):
        super().__init__(f"oracle status {status}: {msg}")
        self.status = status


def score_with_shift(seq1, seq2, matrix, max_shift, shift_penalty=0):
    """ShiftedScorer.scoreWithShift(seq1, seq2) -> (score, shift).  seq = str or u8 codes."""
    a = encode(seq1) if isinstance(seq1, str) else np.ascontiguousarray(seq1, dtype=np.uint8)
    b = encode(seq2) if isinstance(seq2, str) else np.ascontiguousarray(seq2, dtype=np.uint8)
    M = np.ascontiguousarray(matrix, dtype=np.int32).reshape(-1)
    sh, st = C.c_int(0), C.c_int(0)
    sc = lib().hmko_score_with_shift(_p(a, C.c_uint8), len(a), _p(b, C.c_uint8), len(b), _p(M, C.c_int32),
                                     int(max_shift), int(shift_penalty), C.byref(sh), C.byref(st))
    if st.value:
        raise OracleError(st.value, "Shift too big")
    return int(sc), int(sh.value)


def pair_cells(l1, l2, X):
    return int(lib().hmko_pair_cells(l1, l2, X))


def pair_shifts(l1, l2, X):
    return int(lib().hmko_pair_shifts(l1, l2, X))


def sort_order_size(residues, offsets, abundance) -> np.ndarray:
    n = len(abundance)
    perm = np.empty(n, dtype=np.int32)
    lib().hmko_sort_order_size(n, _p(residues, C.c_uint8), _p(offsets, C.c_int32), _p(abundance, C.c_int32),
                               _p(perm, C.c_int32))
    return perm


def default_params(offsets):
    n = len(offsets) - 1
    t, x, k = C.c_int32(), C.c_int32(), C.c_int32()
    lib().hmko_default_params(n, _p(offsets, C.c_int32), C.byref(t), C.byref(x), C.byref(k))
    return int(t.value), int(x.value), int(k.value)


def check_max_shift(offsets, max_shift):
    return int(lib().hmko_check_max_shift(len(offsets) - 1, _p(offsets, C.c_int32), int(max_shift)))


@dataclass
class GreedyResult:
    status: int
    cluster_id: np.ndarray
    member_rank: np.ndarray
    result_order: np.ndarray
    n_multi: int
    counters: dict


def greedy_cluster(residues, offsets, abundance, matrix, threshold, max_shift, shift_penalty, max_clusters,
                   nthreads=1, max_p1_steps=0, max_p2_queries=0) -> GreedyResult:
    residues = np.ascontiguousarray(residues, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.int32)
    abundance = np.ascontiguousarray(abundance, dtype=np.int32)
    M = np.ascontiguousarray(matrix, dtype=np.int32).reshape(-1)
    n = len(abundance)
    cid = np.empty(n, dtype=np.int32)
    rank = np.empty(n, dtype=np.int32)
    order = np.empty(n, dtype=np.int32)
    nres, nmulti = C.c_int32(0), C.c_int32(0)
    ctr = Counters()
    rc = lib().hmko_greedy_cluster_bounded(
        n, _p(residues, C.c_uint8), _p(offsets, C.c_int32), _p(abundance, C.c_int32), _p(M, C.c_int32),
        int(threshold), int(max_shift), int(shift_penalty), int(max_clusters), int(nthreads),
        int(max_p1_steps), int(max_p2_queries),
        _p(cid, C.c_int32), _p(rank, C.c_int32), _p(order, C.c_int32), C.byref(nres), C.byref(nmulti), C.byref(ctr))
    return GreedyResult(rc, cid, rank, order[:nres.value].copy(), int(nmulti.value), ctr.as_dict())


@dataclass
class ClinkageResult:
    status: int
    cluster_id: np.ndarray
    member_rank: np.ndarray
    result_order: np.ndarray
    nearest_searches: int


def clinkage_cluster(residues, offsets, abundance, matrix, threshold, max_shift, shift_penalty) -> ClinkageResult:
    """ClinkageSequenceClusterer.cluster (clinkage_oracle.c); sequences in the caller's order"""
    residues = np.ascontiguousarray(residues, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.int32)
    abundance = np.ascontiguousarray(abundance, dtype=np.int32)
    M = np.ascontiguousarray(matrix, dtype=np.int32).reshape(-1)
    n = len(abundance)
    cid = np.zeros(max(n, 1), dtype=np.int32)
    rank = np.zeros(max(n, 1), dtype=np.int32)
    order = np.zeros(max(n, 1), dtype=np.int32)
    nres, ns = C.c_int32(0), C.c_int64(0)
    rc = lib().hmko_clinkage_cluster(n, _p(residues, C.c_uint8), _p(offsets, C.c_int32), _p(abundance, C.c_int32),
                                     _p(M, C.c_int32), int(threshold), int(max_shift), int(shift_penalty),
                                     _p(cid, C.c_int32), _p(rank, C.c_int32), _p(order, C.c_int32), C.byref(nres), C.byref(ns))
    return ClinkageResult(rc, cid[:n], rank[:n], order[:nres.value].copy(), int(ns.value))


def load_fasta(path: str):
    """-> (strings list, residues, offsets, abundance) in first-occurrence order."""
    f = _Fasta()
    err = C.create_string_buffer(256)
    rc = lib().hmko_load_fasta(path.encode(), C.byref(f), err, 256)
    if rc:
        raise OracleError(rc, err.value.decode())
    try:
        n = f.n
        offs = np.ctypeslib.as_array(f.offsets, shape=(n + 1,)).copy()
        total = int(offs[-1])
        res = np.ctypeslib.as_array(f.residues, shape=(max(total, 1),))[:total].copy()
        ab = np.ctypeslib.as_array(f.abundance, shape=(max(n, 1),))[:n].copy()
        text = C.string_at(f.text, total).decode("latin-1")
        strs = [text[offs[i]:offs[i + 1]] for i in range(n)]
    finally:
        lib().hmko_fasta_free(C.byref(f))
    return strs, res, offs, ab
