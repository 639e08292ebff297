import importlib as _il
import sys as _sys

bias_act = _il.import_module("3dhumangan_b200.ops.bias_act")
upfirdn2d = _il.import_module("3dhumangan_b200.ops.upfirdn2d")
_sys.modules[__name__ + ".bias_act"] = bias_act
_sys.modules[__name__ + ".upfirdn2d"] = upfirdn2d
