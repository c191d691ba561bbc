"""Drop-in for PartGS's ``simple_knn`` extension (submodules/simple-knn)."""
from . import _C  # noqa: F401
