import os
import sys

for _p in sys.path:
    _cand = os.path.join(_p, "lib", "components")
    if os.path.isdir(_cand) and os.path.abspath(_cand) != os.path.dirname(os.path.abspath(__file__)) and _cand not in __path__:
        __path__.append(_cand)
