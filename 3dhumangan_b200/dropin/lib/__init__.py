"""Drop-in `lib` package: hot-path subpackages come from 3dhumangan_b200, the rest from the reference."""
import os
import sys

# let `lib.data`, `lib.trainers`, ... resolve to a reference checkout that is further down sys.path
for _p in sys.path:
    _cand = os.path.join(_p, "lib")
    if os.path.isdir(_cand) and os.path.abspath(_cand) != os.path.dirname(os.path.abspath(__file__)):
        if os.path.exists(os.path.join(_cand, "trainers")) and _cand not in __path__:
            __path__.append(_cand)
