"""Stub: bioframe is imported by skgenome.gary/merge but unused on the hot path."""


def _unavailable(*a, **k):
    raise NotImplementedError("bioframe stub")


overlap = cluster = merge = closest = _unavailable
