"""Import shim for Flask (reference viewer.py:4, pulled in by the reference's cli.py:12).

TEST INFRASTRUCTURE ONLY.  The viewer is out of scope; the shim only lets the reference's `cli` module be
imported so that tests can compare the option surface of its bam2ec / bam2emase commands with ours."""


class Flask(object):
    def __init__(self, *args, **kwargs):
        pass

    def route(self, *args, **kwargs):
        return lambda fn: fn

    def run(self, *args, **kwargs):
        raise NotImplementedError("Flask is not available in this image")


def jsonify(*args, **kwargs):
    raise NotImplementedError("Flask is not available in this image")
