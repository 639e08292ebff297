import importlib as _il

Map3DGenerator = _il.import_module("3dhumangan_b200.modules.generator").Map3DGenerator
