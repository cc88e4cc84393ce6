"""Import shim for the hyphenated package directory ``icp-4dradar_b200/``.

``from icp4r_loader import pkg`` registers it in ``sys.modules`` as ``icp4dradar_b200`` so that
``import icp4dradar_b200.synth`` etc. work afterwards.
"""
import importlib.util
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
_NAME = "icp4dradar_b200"


def _load():
    if _NAME in sys.modules:
        return sys.modules[_NAME]
    d = os.path.join(_ROOT, "icp-4dradar_b200")
    spec = importlib.util.spec_from_file_location(_NAME, os.path.join(d, "__init__.py"), submodule_search_locations=[d])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_NAME] = mod
    spec.loader.exec_module(mod)
    return mod


pkg = _load()
