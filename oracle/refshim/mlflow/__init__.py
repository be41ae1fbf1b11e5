"""Container-only stand-in for ``mlflow`` logging calls (TEST INFRASTRUCTURE). No-ops."""

from types import SimpleNamespace

from . import pyfunc  # noqa: F401

_RUN = None


def log_params(*_, **__):
    return None


def log_metrics(*_, **__):
    return None


def start_run(*_, **__):
    global _RUN
    _RUN = SimpleNamespace(info=SimpleNamespace(run_id="shim"))
    return _RUN


def end_run(*_, **__):
    global _RUN
    _RUN = None


def delete_run(*_, **__):
    return None


def active_run():
    return _RUN
