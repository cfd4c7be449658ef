"""No-op ``matplotlib.pyplot`` (see the package docstring)."""


class _Null:
    def __call__(self, *a, **k):
        return self

    def __getattr__(self, name):
        return self

    def __iter__(self):
        return iter(())


_null = _Null()


def __getattr__(name):
    return _null
