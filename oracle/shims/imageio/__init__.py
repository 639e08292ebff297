"""Test-only stub of imageio."""
def mimwrite(*a, **k):
    pass
def imwrite(*a, **k):
    pass
