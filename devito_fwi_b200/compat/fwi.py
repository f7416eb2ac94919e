"""Drop-in for the reference's top-level ``fwi`` module (minimize.py:4 does ``from fwi import fwi_loss``)."""
from devito_fwi_b200.fwi import *  # noqa: F401,F403
from devito_fwi_b200.fwi import (fm_single, fm_multi, fwi_obj_single, fwi_obj_multi, fwi_loss,  # noqa: F401
                                 fix_source_illumination, resample, Filter, least_square)
