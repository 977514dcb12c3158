from .rrng import RRNG  # noqa: F401
