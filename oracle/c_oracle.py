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


_lib = None


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.ec_oracle_build.restype = ctypes.c_int
        _lib.ec_oracle_build.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                         ctypes.c_int, ctypes.POINTER(_Result)]
        _lib.ec_oracle_free.argtypes = [ctypes.POINTER(_Result)]
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
