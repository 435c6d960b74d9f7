"""Import-only stub, see matplotlib/__init__.py."""


def __getattr__(name):
    raise NotImplementedError(f"matplotlib.pyplot.{name}: plotting is outside the hot path; shim stub")
