"""Import shim for PyTables (reference Sparse3DMatrix.py:4, AlignmentPropertyMatrix.py:3).

TEST INFRASTRUCTURE ONLY.  PyTables is not installable in this image, so no .h5 file can be written here.
What CAN be pinned is everything the writers hand to PyTables: with recording switched on
(`tables.start_recording()`), open_file() returns an object that notes every create_group / create_carray /
set_node_attr call - node path, title, array dtype / shape / values, filter settings - in call order.  The
unmodified reference's APM.save() and alntools_b200.emase.save_emase() are both run against it and must leave
the same record (oracle/make_golden_emase.py, tests/test_emase_calls.py).  Without recording, open_file()
raises as before: nothing pretends to write a file.
"""
import numpy as np

_RECORDING = None   # {filename: [event, ...]} while recording


def start_recording():
    global _RECORDING
    _RECORDING = {}
    return _RECORDING


def stop_recording():
    global _RECORDING
    rec, _RECORDING = _RECORDING, None
    return rec


class Filters(object):
    def __init__(self, complevel=0, complib="zlib", **kwargs):
        self.complevel = complevel
        self.complib = complib
        self.extra = dict(kwargs)

    def describe(self):
        return {"complevel": int(self.complevel), "complib": str(self.complib), "extra": sorted(self.extra)}


class NoSuchNodeError(Exception):
    pass


class _Node(object):
    def __init__(self, path):
        self._v_pathname = path


def _path(where, name):
    base = where._v_pathname if isinstance(where, _Node) else str(where)
    return (base.rstrip("/") + "/" + name) if name else base


def _plain(value):
    """Attribute values in a comparable, JSON-friendly form."""
    if isinstance(value, (list, tuple)):
        return [_plain(v) for v in value]
    if isinstance(value, np.generic):
        return value.item()
    if isinstance(value, np.ndarray):
        return {"ndarray": value.tolist(), "dtype": str(value.dtype)}
    return value


class _RecordingFile(object):
    def __init__(self, events, filename, mode, title):
        self.events = events
        self.root = _Node("/")
        events.append({"op": "open", "mode": mode, "title": title})

    def set_node_attr(self, where, attrname, attrvalue):
        self.events.append({"op": "attr", "node": _path(where, ""), "name": attrname, "value": _plain(attrvalue),
                            "pytype": type(attrvalue).__name__})

    def create_group(self, where, name, title=""):
        p = _path(where, name)
        self.events.append({"op": "group", "node": p, "title": title})
        return _Node(p)

    def create_carray(self, where, name, atom=None, shape=None, title="", filters=None, obj=None, **kwargs):
        p = _path(where, name)
        arr = np.asarray(obj)
        self.events.append({"op": "carray", "node": p, "title": title, "dtype": str(arr.dtype), "shape": list(arr.shape),
                            "filters": filters.describe() if filters is not None else None, "array": arr,
                            "extra": sorted(kwargs)})
        return _Node(p)

    def flush(self):
        pass

    def close(self):
        self.events.append({"op": "close"})


def open_file(filename, mode="r", title="", **kwargs):
    if _RECORDING is None:
        raise NotImplementedError("PyTables is not available in this image")
    if mode not in ("w", "a"):
        raise NotImplementedError("the recording shim only writes")
    events = _RECORDING.setdefault(filename, [])
    if mode == "w":
        del events[:]
    return _RecordingFile(events, filename, mode, title)
