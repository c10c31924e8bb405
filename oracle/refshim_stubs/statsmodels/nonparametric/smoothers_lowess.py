"""Stub: LOWESS is opt-in (bias_smoother='loess') and out of scope."""


def lowess(*a, **k):
    raise NotImplementedError("statsmodels stub")
