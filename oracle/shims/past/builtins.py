"""Import shim for `from past.builtins import xrange` (reference utils.py:16). TEST INFRASTRUCTURE ONLY."""
xrange = range
