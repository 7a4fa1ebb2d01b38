"""Import alias: ``import equss_b200`` loads the package that lives in the directory
``expand-and-quantize-for-unsupervised-semantic-segmentation_b200/`` (a name Python cannot import
directly because of the hyphens).  Submodules resolve normally: ``equss_b200.quantizer`` etc."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)),
                         "expand-and-quantize-for-unsupervised-semantic-segmentation_b200")
_spec = _ilu.spec_from_file_location("equss_b200", _os.path.join(_PKG_DIR, "__init__.py"),
                                     submodule_search_locations=[_PKG_DIR])
_mod = _ilu.module_from_spec(_spec)
_sys.modules["equss_b200"] = _mod
_spec.loader.exec_module(_mod)
