"""Import-only stub (test infrastructure): the reference imports pyvista at module top."""


class PolyData:  # noqa: D101
    pass


class UnstructuredGrid:  # noqa: D101
    pass


class CellType:
    QUAD = 9


def get_reader(*a, **k):
    raise RuntimeError("pyvista stub")


def start_xvfb(*a, **k):
    pass
