import importlib as _il

UNetDiscriminator = _il.import_module("3dhumangan_b200.modules.discriminator").UNetDiscriminator
