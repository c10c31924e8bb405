"""Stub: only `as_handle` is used by skgenome.tabio."""
import contextlib


@contextlib.contextmanager
def as_handle(handleish, mode="r", **kwargs):
    if isinstance(handleish, (str, bytes)) or hasattr(handleish, "__fspath__"):
        with open(handleish, mode, **kwargs) as fp:
            yield fp
    else:
        yield handleish
