"""Host-side mirrors of the reference's module surfaces, built on the C ABI (no CPU fallback)."""
