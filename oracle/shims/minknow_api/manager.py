"""Import shim (test infrastructure)."""


class Manager:
    def __init__(self, *args, **kwargs):
        pass


class FlowCellPosition:
    pass
