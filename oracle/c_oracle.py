"""ctypes loader of the C oracle (oracle/ec_oracle.c).  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(BUILD_DIR, "libecoracle.so")


def build(force=False):
    src = os.path.join(HERE, "ec_oracle.c")
    if not force and os.path.isfile(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(src):
        return LIB
    os.makedirs(BUILD_DIR, exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", LIB, src])
    return LIB


class _Result(ctypes.Structure):
    _fields_ = [("n_ec", ctypes.c_int64), ("nnz", ctypes.c_int64),
                ("indptr", ctypes.POINTER(ctypes.c_int32)), ("indices", ctypes.POINTER(ctypes.c_int32)),
                ("data", ctypes.POINTER(ctypes.c_int32)), ("counts", ctypes.POINTER(ctypes.c_int32)),
                ("n_reads", ctypes.c_int64)]


class _CellsResult(ctypes.Structure):
    _P = ctypes.POINTER(ctypes.c_int32)
    _fields_ = [("n_ec", ctypes.c_int64), ("nnz_a", ctypes.c_int64), ("a_indptr", _P), ("a_indices", _P), ("a_data", _P),
                ("n_cells", ctypes.c_int64), ("nnz_n", ctypes.c_int64), ("n_indptr", _P), ("n_indices", _P),
                ("n_data", _P), ("cell_order", _P), ("n_reads", ctypes.c_int64)]


_lib = None


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.ec_oracle_build.restype = ctypes.c_int
        _lib.ec_oracle_build.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                         ctypes.c_int, ctypes.POINTER(_Result)]
        _lib.ec_oracle_free.argtypes = [ctypes.POINTER(_Result)]
        pp = ctypes.POINTER(ctypes.c_void_p)
        _lib.ec_oracle_build_cells.restype = ctypes.c_int
        _lib.ec_oracle_build_cells.argtypes = [ctypes.c_int, pp, pp, pp, pp, ctypes.POINTER(ctypes.c_int64),
                                               ctypes.POINTER(ctypes.c_int32), ctypes.c_int64,
                                               ctypes.POINTER(_CellsResult)]
        _lib.ec_oracle_cells_free.argtypes = [ctypes.POINTER(_CellsResult)]
    return _lib


def ec_from_columns(read_group, target_idx, hap_idx, drop_last=False):
    """Same contract as oracle.ec_oracle.ec_from_columns, plus n_reads."""
    lib = _load()
    rg = np.ascontiguousarray(read_group, dtype=np.int32)
    tg = np.ascontiguousarray(target_idx, dtype=np.int32)
    hp = np.ascontiguousarray(hap_idx, dtype=np.int32)
    res = _Result()
    lib.ec_oracle_build(rg.ctypes.data, tg.ctypes.data, hp.ctypes.data, len(rg), 1 if drop_last else 0,
                        ctypes.byref(res))

    def arr(p, n):
        return np.ctypeslib.as_array(p, shape=(n,)).copy() if n else np.zeros(0, dtype=np.int32)

    out = (arr(res.indptr, res.n_ec + 1), arr(res.indices, res.nnz), arr(res.data, res.nnz),
           arr(res.counts, res.n_ec), int(res.n_reads))
    lib.ec_oracle_free(ctypes.byref(res))
    return out


def ec_from_columns_cells(pushes, minimum_count):
    """Same contract as oracle.ec_oracle.ec_from_columns_cells (pushes: list, in file order, of
    (read_group, target_idx, hap_idx, cell_idx, drop_last_group)), in C: the per-cell path at sizes the
    Python statement cannot finish."""
    lib = _load()
    n = len(pushes)
    keep = []
    cols = [(ctypes.c_void_p * n)() for _ in range(4)]
    rows = (ctypes.c_int64 * n)()
    drop = (ctypes.c_int32 * n)()
    for i, (rg, tg, hp, cell, drop_last) in enumerate(pushes):
        arrs = [np.ascontiguousarray(a, dtype=np.int32) for a in (rg, tg, hp, cell)]
        keep.append(arrs)
        for j in range(4):
            cols[j][i] = arrs[j].ctypes.data
        rows[i] = len(arrs[0])
        drop[i] = 1 if drop_last else 0
    res = _CellsResult()
    rc = lib.ec_oracle_build_cells(n, cols[0], cols[1], cols[2], cols[3], rows, drop, int(minimum_count), ctypes.byref(res))
    if rc == -2:
        raise ValueError('max() arg is an empty sequence')
    if rc != 0:
        raise ValueError("ec_oracle_build_cells: bad input (negative cell id?)")

    def arr(p, k):
        return np.ctypeslib.as_array(p, shape=(k,)).copy() if k else np.zeros(0, dtype=np.int32)

    out = {"a_indptr": arr(res.a_indptr, res.n_ec + 1), "a_indices": arr(res.a_indices, res.nnz_a),
           "a_data": arr(res.a_data, res.nnz_a), "n_indptr": arr(res.n_indptr, res.n_cells + 1),
           "n_indices": arr(res.n_indices, res.nnz_n), "n_data": arr(res.n_data, res.nnz_n),
           "cell_order": arr(res.cell_order, res.n_cells), "n_reads": int(res.n_reads)}
    lib.ec_oracle_cells_free(ctypes.byref(res))
    return out
