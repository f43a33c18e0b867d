"""Import-only stub (test infrastructure)."""


def Fire(*a, **k):
    raise RuntimeError("fire stub")
