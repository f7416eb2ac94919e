"""Drop-in for ``examples.seismic`` (what seismic/inversion/fwi.py:5-6 and the tutorials import)."""
from devito_fwi_b200.model import *  # noqa: F401,F403
from devito_fwi_b200.source import *  # noqa: F401,F403
from devito_fwi_b200.geometry import *  # noqa: F401,F403
from devito_fwi_b200.preset_models import *  # noqa: F401,F403


def plot_velocity(*args, **kwargs):      # plotting is outside the hot path (seismic/plotting.py)
    pass


plot_shotrecord = plot_perturbation = plot_image = plot_velocity
