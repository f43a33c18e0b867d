"""Import-only stub (test infrastructure)."""


class Mesh:  # noqa: D101
    pass


class DataSet:  # noqa: D101
    pass
