"""Test-only stub of `smplx` (absent from this image). No arithmetic on the hot path."""
