"""numpy-in / numpy-out mirror of `malstroem.algorithms` (fill, flow, label) on the B200 library."""
from . import fill, flow, label  # noqa: F401
