"""Stub of matplotlib (not installed here); test infrastructure so the unmodified
reference imports.  Every attribute resolves to a dummy object."""
import sys, types


class _Dummy:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Dummy()

    def __getattr__(self, name):
        return _Dummy()

    def __iter__(self):
        return iter(())


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    return _Dummy
