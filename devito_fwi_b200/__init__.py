"""devito_fwi_b200 -- B200-native drop-in for the FWI-gradient hot path of LongyanU/devito-fwi.

Host-side mirror of the reference's ``seismic`` package surface (Model, sources, geometry,
AcousticWaveSolver) and of ``fwi.py``; all arithmetic of the path runs in hand-written sm_100a
CUDA kernels behind the C ABI of ``libb2fwi.so`` (include/b2fwi.h).
"""
from .grid import Grid, Function, TimeFunction, Constant, norm, mmax, mmin  # noqa: F401
from .model import SeismicModel, Model, initialize_damp, initialize_function  # noqa: F401
from .source import (TimeAxis, PointSource, Receiver, Shot, WaveletSource, RickerSource,  # noqa: F401
                     GaborSource, DGaussSource)
from .geometry import AcquisitionGeometry, setup_geometry, setup_rec_coords  # noqa: F401
from .preset_models import demo_model  # noqa: F401
from .wavesolver import AcousticWaveSolver, PerformanceSummary  # noqa: F401
from . import fwi, dist, resident  # noqa: F401

__version__ = "0.1.0"
