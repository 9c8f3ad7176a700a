"""Import shim for PyTables (reference Sparse3DMatrix.py:4, AlignmentPropertyMatrix.py:3).

TEST INFRASTRUCTURE ONLY. PyTables is not installable in this image, so the emase (.h5)
writer of the reference cannot be executed; only the import has to succeed.
"""


def open_file(*args, **kwargs):
    raise NotImplementedError("PyTables is not available in this image")


class Filters(object):
    def __init__(self, *args, **kwargs):
        pass


class NoSuchNodeError(Exception):
    pass
