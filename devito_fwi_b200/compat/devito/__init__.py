"""Minimal ``devito`` namespace for the names the hot path's callers import
(fwi.py:1 ``from devito import Function``; seismic/inversion/fwi.py:3 ``configuration, Function, norm, mmax, mmin``)."""
from devito_fwi_b200.grid import Grid, Function, TimeFunction, Constant, norm, mmax, mmin  # noqa: F401

configuration = {'log-level': 'INFO'}


def set_log_level(level, comm=None):
    configuration['log-level'] = level
