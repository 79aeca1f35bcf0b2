"""No-op stand-in for the third-party ``deprecation`` package (reference vertical/array/vertical.py:14,23).

TEST INFRASTRUCTURE ONLY: lets the unmodified reference's ``earthkit.meteo.vertical`` import in the build
container when the hybrid-level golden vectors are generated (tests/golden/make_golden.py).
"""


def deprecated(*args, **kwargs):
    def wrap(fn):
        return fn

    return wrap


def fail_if_not_removed(fn):
    return fn
