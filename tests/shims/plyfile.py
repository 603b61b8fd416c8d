"""Stand-in for `plyfile` (absent offline): the reference only uses it in save_ply / load_ply, which the loop test does
not call."""


class PlyElement:
    @staticmethod
    def describe(*a, **k):
        raise NotImplementedError("plyfile is not available offline")


class PlyData:
    def __init__(self, *a, **k):
        raise NotImplementedError("plyfile is not available offline")

    @staticmethod
    def read(*a, **k):
        raise NotImplementedError("plyfile is not available offline")
