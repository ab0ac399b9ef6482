"""Locate and load the REFERENCE's copy of a module that compat/ shadows (nothing is copied: the file is executed
from where it lies on sys.path, under a private module name)."""
import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))


def find(modname):
    """Path of `<entry>/<mod/name>.py` in the first sys.path entry other than compat/ that has it, or None."""
    rel = os.path.join(*modname.split(".")) + ".py"
    for entry in sys.path:
        root = os.path.abspath(entry or os.getcwd())
        if root == _HERE:
            continue
        cand = os.path.join(root, rel)
        if os.path.isfile(cand) and not os.path.abspath(cand).startswith(_HERE + os.sep):
            return cand
    return None


def load(modname):
    """The reference's module `modname` (e.g. 'src.functions') or None when the reference tree is not on sys.path.
    Import errors of the reference module itself (a missing third-party wheel) propagate as ImportError."""
    key = "_f4l_upstream." + modname
    if key in sys.modules:
        return sys.modules[key]
    path = find(modname)
    if path is None:
        return None
    spec = importlib.util.spec_from_file_location(key, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[key] = mod
    try:
        spec.loader.exec_module(mod)
    except BaseException:
        del sys.modules[key]
        raise
    return mod


def reexport(modname, into, skip=()):
    """Copy every public name of the reference module into the namespace `into` (names in `skip` stay ours)."""
    up = load(modname)
    if up is None:
        return None
    for k, v in vars(up).items():
        if not k.startswith("__") and k not in skip:
            into.setdefault(k, v)
    return up
