class Digraph:  # pragma: no cover
    def __init__(self, *a, **k):
        raise NotImplementedError("graphviz is not available")
