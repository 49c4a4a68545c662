def __getattr__(name):
    def _noop(*a, **k):
        raise RuntimeError("matplotlib stub: plotting is not available")
    return _noop
