"""ctypes binding of libecb200.so (C ABI declared in include/ecb200.h).

The library is built in-tree by `__graft_entry__.build()` (nvcc, sm_100a).  There is no CPU
implementation behind this module: if the shared object is missing or no B200 is visible the
constructors raise.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libecb200.so")

ECB_OPT_RESULT_ON_DEVICE = 1
ECB_OPT_TABLE_SLOTS = 2
ECB_OPT_PAIR_SLOTS = 3
ECB_OPT_GRID_CTAS = 4
ECB_OPT_HOT_CACHE = 5
ECB_OPT_VERIFY_KEYS = 6
ECB_OPT_CHUNK_LEN = 7
ECB_OPT_PAGEABLE_RESULTS = 8

ECB_ERR_EMPTY = -5

_i32p = ctypes.POINTER(ctypes.c_int32)


class EcbResult(ctypes.Structure):
    _fields_ = [
        ("n_ec", ctypes.c_int64), ("nnz_a", ctypes.c_int64),
        ("a_indptr", _i32p), ("a_indices", _i32p), ("a_data", _i32p),
        ("n_samples", ctypes.c_int64), ("nnz_n", ctypes.c_int64),
        ("n_indptr", _i32p), ("n_indices", _i32p), ("n_data", _i32p),
        ("cell_order", _i32p),
        ("n_reads", ctypes.c_int64), ("n_alignments", ctypes.c_int64),
    ]


class EcbExport(ctypes.Structure):
    _fields_ = [
        ("n_ec", ctypes.c_int64), ("n_rows", ctypes.c_int64),
        ("meta", ctypes.POINTER(ctypes.c_int64)), ("rows", _i32p),
        ("part_ec_counts", ctypes.POINTER(ctypes.c_int64)), ("part_row_counts", ctypes.POINTER(ctypes.c_int64)),
        ("min_base", ctypes.c_int64), ("max_end", ctypes.c_int64),
    ]


class EcbStats(ctypes.Structure):
    _fields_ = [
        ("group_ms", ctypes.c_double), ("harvest_ms", ctypes.c_double),
        ("push_ms", ctypes.c_double), ("finalize_ms", ctypes.c_double),
        ("kernel_launches", ctypes.c_int64), ("table_slots", ctypes.c_int64),
        ("table_used", ctypes.c_int64), ("table_grows", ctypes.c_int64),
        ("overflow_reads", ctypes.c_int64), ("h2d_bytes", ctypes.c_int64),
        ("d2h_bytes", ctypes.c_int64), ("row_entries", ctypes.c_int64),
    ]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


class EcbSlice(ctypes.Structure):
    _i32p = ctypes.POINTER(ctypes.c_int32)
    _fields_ = [("id_base", ctypes.c_int64), ("n_ec", ctypes.c_int64), ("nnz", ctypes.c_int64),
                ("a_indptr", _i32p), ("a_indices", _i32p), ("a_data", _i32p), ("n_data", _i32p)]


# every symbol include/ecb200.h declares, with its ctypes signature
SIGNATURES = {
    "ecb_version": (ctypes.c_int, []),
    "ecb_create": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                  ctypes.c_int, ctypes.c_int64]),
    "ecb_set_option": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int64]),
    "ecb_set_stream": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "ecb_arena_create": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)]),
    "ecb_arena_open_peer": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)]),
    "ecb_arena_reset": (ctypes.c_int, [ctypes.c_void_p]),
    "ecb_rebase": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64]),
    "ecb_order_dispatch": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p), ctypes.c_int64,
                                          ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64)]),
    "ecb_order_build": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.POINTER(EcbSlice)]),
    "ecb_export_to_arenas": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p), ctypes.c_int64,
                                            ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64)]),
    "ecb_import_arena": (ctypes.c_int, [ctypes.c_void_p]),
    "ecb_push": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_int]),
    "ecb_finalize": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(EcbResult)]),
    "ecb_reset": (ctypes.c_int, [ctypes.c_void_p]),
    "ecb_get_stats": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(EcbStats)]),
    "ecb_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "ecb_last_error": (ctypes.c_char_p, [ctypes.c_void_p]),
    "ecb_export_partition": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(EcbExport)]),
    "ecb_import_entries": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64),
                                          ctypes.c_int]),
    "ecb_global_mark": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64]),
    "ecb_global_count": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                        ctypes.POINTER(ctypes.c_int64)]),
    "ecb_global_lens": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "ecb_global_indptr": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                         ctypes.POINTER(ctypes.c_int64)]),
    "ecb_global_rows": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
}

_lib = None


class EcbError(RuntimeError):
    def __init__(self, code, message):
        RuntimeError.__init__(self, "libecb200 error %d: %s" % (code, message))
        self.code = code


def load_library(path=None):
    """dlopen libecb200.so and attach signatures.  Raises if the extension has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("ECB200_LIB") or LIB_PATH
    if not os.path.isfile(p):
        raise ImportError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback for the EC build)" % p)
    lib = ctypes.CDLL(p)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    if path is None:
        _lib = lib
    return lib


def _ptr(x):
    """Address of a numpy array / torch tensor / raw int; keeps nothing alive."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if isinstance(x, np.ndarray):
        if x.dtype != np.int32 or not x.flags["C_CONTIGUOUS"]:
            raise TypeError("columns must be C-contiguous int32 arrays")
        return x.ctypes.data
    if hasattr(x, "data_ptr"):  # torch tensor (host pinned or device); plumbing only
        import torch
        if x.dtype != torch.int32 or not x.is_contiguous():
            raise TypeError("columns must be contiguous int32 tensors")
        return x.data_ptr()
    raise TypeError("unsupported column type %r" % type(x))


class _CudaArray(object):
    """Zero-copy hand-off of library-owned device memory to torch (CUDA array interface)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 2}


def _device_view(ptr, shape, typestr, dtype, device):
    import torch
    if shape[0] == 0 or not ptr:
        return torch.empty(shape, dtype=dtype, device=device)
    return torch.as_tensor(_CudaArray(ptr, shape, typestr), device=device)


def _on_device(x):
    return bool(getattr(x, "is_cuda", False))


class EcBuilder(object):
    """One context = one GPU.  push() columns, finalize() -> dict of numpy int32 arrays."""

    def __init__(self, n_targets, n_haps, with_cells=False, alignments_hint=0, device=0, **options):
        self._lib = load_library()
        self._ctx = ctypes.c_void_p()
        rc = self._lib.ecb_create(ctypes.byref(self._ctx), device, int(n_targets), int(n_haps),
                                  1 if with_cells else 0, int(alignments_hint))
        if rc != 0:
            raise EcbError(rc, self._lib.ecb_last_error(None).decode())
        self.with_cells = bool(with_cells)
        self._result_on_device = False
        for key, value in options.items():
            self.set_option(key, value)

    _OPTIONS = {"result_on_device": ECB_OPT_RESULT_ON_DEVICE, "table_slots": ECB_OPT_TABLE_SLOTS,
                "pair_slots": ECB_OPT_PAIR_SLOTS, "grid_ctas": ECB_OPT_GRID_CTAS,
                "hot_cache": ECB_OPT_HOT_CACHE, "verify_keys": ECB_OPT_VERIFY_KEYS,
                "chunk_len": ECB_OPT_CHUNK_LEN, "pageable_results": ECB_OPT_PAGEABLE_RESULTS}

    def _check(self, rc):
        if rc != 0:
            raise EcbError(rc, self._lib.ecb_last_error(self._ctx).decode())

    def set_option(self, name, value):
        self._check(self._lib.ecb_set_option(self._ctx, self._OPTIONS[name], int(value)))
        if name == "result_on_device":
            self._result_on_device = bool(value)

    def set_stream(self, cuda_stream):
        """cuda_stream: a cudaStream_t as int (e.g. torch.cuda.current_stream().cuda_stream), or None for
        the library's own stream.  torch reports its default stream as handle 0, which the C ABI reads
        as "own stream"; it is passed as cudaStreamLegacy (0x1), the explicit name of the same stream."""
        if cuda_stream is not None and int(cuda_stream) == 0:
            cuda_stream = 1
        self._check(self._lib.ecb_set_stream(self._ctx, cuda_stream))

    def push(self, read_group, target_idx, hap_idx, cell_idx=None, order_base=0, drop_last_group=False,
             n=None):
        if n is None:
            n = int(read_group.shape[0])
        dev = _on_device(read_group)
        self._check(self._lib.ecb_push(self._ctx, _ptr(read_group), _ptr(target_idx), _ptr(hap_idx),
                                       _ptr(cell_idx), n, int(order_base), 1 if drop_last_group else 0,
                                       1 if dev else 0))

    def finalize_raw(self, min_cell_count=0):
        res = EcbResult()
        self._check(self._lib.ecb_finalize(self._ctx, int(min_cell_count), ctypes.byref(res)))
        return res

    def finalize(self, min_cell_count=0, copy=True):
        """Returns dict(a_indptr, a_indices, a_data, n_indptr, n_indices, n_data[, cell_order]) as
        numpy int32 arrays plus scalar sizes.  copy=False returns VIEWS of the library's result buffers
        (valid until the next finalize / reset / close): the EC file can be written straight from them."""
        if self._result_on_device:
            raise RuntimeError("finalize() needs host results; use finalize_raw() with result_on_device")
        res = self.finalize_raw(min_cell_count)

        def arr(ptr, n):
            if n == 0:
                return np.zeros(0, dtype=np.int32)
            view = np.ctypeslib.as_array(ptr, shape=(n,))
            return view.copy() if copy else view

        out = {
            "n_ec": res.n_ec, "nnz_a": res.nnz_a, "n_samples": res.n_samples, "nnz_n": res.nnz_n,
            "n_reads": res.n_reads, "n_alignments": res.n_alignments,
            "a_indptr": arr(res.a_indptr, res.n_ec + 1),
            "a_indices": arr(res.a_indices, res.nnz_a),
            "a_data": arr(res.a_data, res.nnz_a),
            "n_indptr": arr(res.n_indptr, res.n_samples + 1),
            "n_indices": arr(res.n_indices, res.nnz_n),
            "n_data": arr(res.n_data, res.nnz_n),
        }
        if self.with_cells:
            out["cell_order"] = arr(res.cell_order, res.n_samples)
        return out

    def reset(self):
        self._check(self._lib.ecb_reset(self._ctx))

    # ---- multi-GPU exchange primitives (see alntools_b200/multi_gpu.py) -------------------------
    def export_partition(self, world):
        """-> (meta int64[n_ec, 5], rows int32[n_rows, 2]) torch device tensors (copies), per-owner
        EC / row counts (lists) and the pushed order-key range."""
        import torch
        exp = EcbExport()
        self._check(self._lib.ecb_export_partition(self._ctx, int(world), ctypes.byref(exp)))
        dev = torch.device("cuda", torch.cuda.current_device())
        # copies into torch-owned memory: the collectives then only ever see torch allocations, and the
        # library is free to reuse its export buffers
        meta = _device_view(ctypes.cast(exp.meta, ctypes.c_void_p).value, (exp.n_ec, 5), "<i8", torch.int64, dev).clone()
        rows = _device_view(ctypes.cast(exp.rows, ctypes.c_void_p).value, (exp.n_rows, 2), "<i4", torch.int32, dev).clone()
        ec_counts = [int(exp.part_ec_counts[i]) for i in range(world)]
        row_counts = [int(exp.part_row_counts[i]) for i in range(world)]
        return meta, rows, ec_counts, row_counts, int(exp.min_base), int(exp.max_end)

    def import_entries(self, meta, rows, ec_counts, row_counts):
        n = len(ec_counts)
        a = (ctypes.c_int64 * n)(*ec_counts)
        b = (ctypes.c_int64 * n)(*row_counts)
        self._check(self._lib.ecb_import_entries(self._ctx, meta.data_ptr(), rows.data_ptr(), a, b, n))

    # ---- exchange over peer memory (CUDA IPC arenas, see include/ecb200.h) -------------------------
    def arena_create(self, cap_records):
        """-> (64-byte IPC handle, base address) of this OWNER context's arena (cap_records records of 32 bytes)."""
        handle = ctypes.create_string_buffer(64)
        base = ctypes.c_void_p()
        self._check(self._lib.ecb_arena_create(self._ctx, int(cap_records), handle, ctypes.byref(base)))
        return handle.raw, int(base.value)

    def arena_open_peer(self, handle):
        base = ctypes.c_void_p()
        self._check(self._lib.ecb_arena_open_peer(self._ctx, ctypes.create_string_buffer(handle, 64), ctypes.byref(base)))
        return int(base.value)

    def arena_reset(self):
        self._check(self._lib.ecb_arena_reset(self._ctx))

    def rebase(self, delta):
        """Shift the order key of everything pushed so far by `delta` (see ecb_rebase in include/ecb200.h)."""
        self._check(self._lib.ecb_rebase(self._ctx, int(delta)))

    def export_to_arenas(self, bases, cap_records):
        """Dispatch 1: every EC of this LOCAL context as {key, first, count} into the arena of its owner rank.
        -> (min_base, max_end) of the positions pushed here."""
        arr = (ctypes.c_void_p * len(bases))(*bases)
        lo, hi = ctypes.c_int64(), ctypes.c_int64()
        self._check(self._lib.ecb_export_to_arenas(self._ctx, len(bases), arr, int(cap_records), ctypes.byref(lo), ctypes.byref(hi)))
        return int(lo.value), int(hi.value)

    def import_arena(self):
        self._check(self._lib.ecb_import_arena(self._ctx))

    def order_dispatch(self, bases, cap_records, shard_lo, shard_hi):
        """Dispatch 2: every merged EC of this OWNER context as {key, position inside its shard, count} to the rank
        whose shard of the read order holds its first occurrence (shard_lo / shard_hi: one entry per rank)."""
        arr = (ctypes.c_void_p * len(bases))(*bases)
        lo = (ctypes.c_int64 * len(bases))(*[int(x) for x in shard_lo])
        hi = (ctypes.c_int64 * len(bases))(*[int(x) for x in shard_hi])
        self._check(self._lib.ecb_order_dispatch(self._ctx, len(bases), arr, int(cap_records), lo, hi))

    def order_build(self, local, shard_lo, shard_hi):
        """-> dict(n_ec, nnz, a_indptr, a_indices, a_data, n_data): the ECs whose first occurrence lies in this
        rank's shard, in id order, rows taken from `local` (the builder that holds this rank's reads), as torch
        views of library memory (valid until the next call on this context)."""
        import torch
        sl = EcbSlice()
        self._check(self._lib.ecb_order_build(self._ctx, local._ctx, int(shard_lo), int(shard_hi), ctypes.byref(sl)))
        dev = torch.device("cuda", torch.cuda.current_device())
        view = lambda p, n: _device_view(ctypes.cast(p, ctypes.c_void_p).value, (n,), "<i4", torch.int32, dev)
        return {"n_ec": int(sl.n_ec), "nnz": int(sl.nnz), "a_indptr": view(sl.a_indptr, sl.n_ec + 1),
                "a_indices": view(sl.a_indices, max(sl.nnz, 1))[:sl.nnz], "a_data": view(sl.a_data, max(sl.nnz, 1))[:sl.nnz],
                "n_data": view(sl.n_data, max(sl.n_ec, 1))[:sl.n_ec]}

    def global_mark(self, min_base, bitmap):
        self._check(self._lib.ecb_global_mark(self._ctx, int(min_base), bitmap.data_ptr(), bitmap.numel()))

    def global_count(self, bitmap):
        total = ctypes.c_int64()
        self._check(self._lib.ecb_global_count(self._ctx, bitmap.data_ptr(), bitmap.numel(), ctypes.byref(total)))
        return int(total.value)

    def global_lens(self, lens, counts):
        self._check(self._lib.ecb_global_lens(self._ctx, lens.data_ptr(), counts.data_ptr()))

    def global_indptr(self, lens):
        nnz = ctypes.c_int64()
        self._check(self._lib.ecb_global_indptr(self._ctx, lens.data_ptr(), lens.numel(), ctypes.byref(nnz)))
        return int(nnz.value)

    def global_rows(self, indptr, indices, data):
        self._check(self._lib.ecb_global_rows(self._ctx, indptr.data_ptr(), indices.data_ptr(), data.data_ptr()))

    def stats(self):
        st = EcbStats()
        self._check(self._lib.ecb_get_stats(self._ctx, ctypes.byref(st)))
        return st.as_dict()

    def close(self):
        if self._ctx:
            self._lib.ecb_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
