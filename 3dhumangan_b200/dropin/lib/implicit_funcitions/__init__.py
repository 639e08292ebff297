import importlib as _il

COORDCONCATSIREN = _il.import_module("3dhumangan_b200.modules.implicit").COORDCONCATSIREN
