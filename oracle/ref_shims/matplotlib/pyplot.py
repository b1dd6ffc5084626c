"""TEST INFRASTRUCTURE ONLY -- import-only stand-in for matplotlib.pyplot."""
def __getattr__(name):
    def _f(*a, **k):
        raise RuntimeError('matplotlib shim: plotting is not available')
    return _f
