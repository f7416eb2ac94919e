"""``from seismic.acoustic import AcousticWaveSolver`` (fwi.py:3)."""
from devito_fwi_b200.wavesolver import AcousticWaveSolver, PerformanceSummary  # noqa: F401
