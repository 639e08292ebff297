import importlib as _il

_c = _il.import_module("3dhumangan_b200.configs")
MAP3DBN, MAP3DBN512, MAP3DBN512L = _c.MAP3DBN, _c.MAP3DBN512, _c.MAP3DBN512L
extract_metadata, get_config = _c.extract_metadata, _c.get_config
