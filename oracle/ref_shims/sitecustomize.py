"""Oracle scaffolding: the reference targets typeguard 2.x; typeguard 4 breaks its
@typechecked functions, so make the decorator the identity (without shadowing the package)."""
try:
    import typeguard

    def _identity(f=None, **k):
        if f is None:
            return lambda g: g
        return f

    typeguard.typechecked = _identity
except Exception:  # pragma: no cover
    pass
