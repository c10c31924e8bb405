"""Stub: FASTA access is out of scope."""


class Fasta:  # pragma: no cover
    def __init__(self, *a, **k):
        raise NotImplementedError("pyfaidx stub")
