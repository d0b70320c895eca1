"""models/mynn.py of the reference is a byte-identical duplicate of models/norm.py; same here."""
from .norm import Norm2d  # noqa: F401
