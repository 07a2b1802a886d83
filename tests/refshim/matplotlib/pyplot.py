"""TEST INFRASTRUCTURE ONLY — every attribute is a callable returning another dummy."""


class _Dummy:
    def __init__(self, *a, **k):
        self.patches = []

    def __call__(self, *a, **k):
        return _Dummy()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Dummy()

    def __iter__(self):
        return iter(())


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    return _Dummy()
