"""Importable alias of the package directory ``extracting-tree-morphology-from-point-clouds_b200/``
(a hyphenated directory name cannot be imported directly).  ``import treemorph_b200`` resolves its
submodules from that directory."""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         "extracting-tree-morphology-from-point-clouds_b200")
__path__.insert(0, _PKG_DIR)

with open(_os.path.join(_PKG_DIR, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_PKG_DIR, "__init__.py"), "exec"))
