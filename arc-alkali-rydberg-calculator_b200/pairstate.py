"""PairStateInteractions -- placeholder, implemented below in this round."""


class PairStateInteractions:
    def __init__(self, *a, **k):
        raise NotImplementedError("PairStateInteractions: not built yet")
