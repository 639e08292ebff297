"""Test-only stub of `pytorch3d` 0.6.2 (absent from this image)."""
__version__ = "0.6.2-stub"
