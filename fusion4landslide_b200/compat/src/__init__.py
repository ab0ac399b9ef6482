# `src` package of the drop-in: modules present here shadow the reference's, every other `src.X` is looked up in the
# reference's own src/ directory (appended to __path__ from the remaining sys.path entries).
import os as _os
import sys as _sys

_here = _os.path.dirname(_os.path.abspath(__file__))
for _e in list(_sys.path):
    _cand = _os.path.join(_os.path.abspath(_e or _os.getcwd()), "src")
    if _os.path.isdir(_cand) and _os.path.abspath(_cand) != _here and _cand not in __path__:
        __path__.append(_cand)
_repo = _os.path.dirname(_os.path.dirname(_os.path.dirname(_here)))
if _repo not in _sys.path:                      # make `import fusion4landslide_b200` work from the mains
    _sys.path.append(_repo)
