"""TEST INFRASTRUCTURE ONLY -- import-only stand-in for matplotlib."""
rcParams = {}
def use(*a, **k):
    pass
