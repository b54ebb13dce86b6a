"""TEST INFRASTRUCTURE ONLY - aslnn.py:14 imports Fabber but never uses it."""


class Fabber:
    def __init__(self, *a, **k):
        raise RuntimeError("Fabber is not available in this environment")
