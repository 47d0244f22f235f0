"""`dfd` — importable alias of the package directory
`deepfake-detection-using-clip-based-siglip-2-vision-transformers_b200/` (whose name is not a Python identifier).

    import dfd
    from dfd import ops, engine, scoring
"""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         "deepfake-detection-using-clip-based-siglip-2-vision-transformers_b200")
if not _os.path.isdir(_PKG_DIR):
    raise ImportError(f"package directory missing: {_PKG_DIR}")
__path__.append(_PKG_DIR)
__version__ = "0.1.0"
PACKAGE_DIR = _PKG_DIR
