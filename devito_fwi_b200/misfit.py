"""Misfit plug-ins with the reference's interface ``misfit(syn, obs) -> (fval, adjoint_source)`` on (nt, nrec) records
(misfit/misfit.py): ``least_square`` (:5-9) and ``qWasserstein`` (:11-104, trans_type='linear', methods '1d' and '2d').

Inside ``fwi_obj_multi`` these objects are recognised and evaluated on the device for all shots of a rank at once
(fwi._objective_resident); called directly they run the same device kernels on the one record given (there is no CPU
implementation in this package). The '2d' method is the back-and-forth optimal-transport solver the reference keeps in
misfit/QW2D/src/fot2d.c and runs per shot as a subprocess over files (misfit/bfm.py:145-193)."""
import ctypes

import numpy as np

from . import _lib
from .fwi import least_square

__all__ = ['least_square', 'qWasserstein', 'Misfit']


class qWasserstein(object):
    """qWasserstein(trans_type='linear', gamma=1.0, method='1d', num_steps=10, step_scale=1.)   [misfit/misfit.py:12-17]"""

    def __init__(self, trans_type='linear', gamma=1.0, method='1d', num_steps=10, step_scale=1.):
        assert method in ['1d', '2d']
        if trans_type != 'linear':
            raise NotImplementedError("only the 'linear' positivity transform (the drivers' choice, "
                                      "marmousi2_fwi.py:131-132) is implemented on the device")
        self.gamma = gamma
        self.method = method
        self.trans_type = trans_type
        self.num_steps = num_steps
        self.step_scale = step_scale

    def __call__(self, f, g):
        """(loss, adjoint source) of one record; float64 adjoint source like the reference's ``grad * d``."""
        import torch
        from .wavesolver import _ptr, _stream
        f = np.ascontiguousarray(f, dtype=np.float32)
        g = np.ascontiguousarray(g, dtype=np.float32)
        if f.ndim != 2 or f.shape != g.shape or f.shape[1] <= 1:
            raise ValueError("Can not use 2d method for 1D input." if self.method == '2d'
                             else "expected two (nt, nrec) records of the same shape")
        nt, nrec = f.shape
        lib = _lib.lib()
        fd, gd = torch.from_numpy(f).cuda(), torch.from_numpy(g).cuda()
        adj = torch.empty_like(fd)
        fval = torch.zeros(1, dtype=torch.float64, device='cuda')
        if self.method == '2d':
            scratch = torch.empty(int(lib.b2fwi_qw2d_scratch_bytes(nt, nrec, 1)), dtype=torch.uint8, device='cuda')
            _lib.check(lib.b2fwi_qw2d_misfit(_ptr(fd), _ptr(gd), None, nt, nrec, 1, ctypes.c_double(float(self.gamma)),
                                             int(self.num_steps), ctypes.c_float(float(self.step_scale)), _ptr(adj),
                                             _ptr(fval), None, _ptr(scratch), _stream()))
        else:
            scratch = torch.empty(int(lib.b2fwi_w1d_scratch_bytes(nt, nrec, 1)), dtype=torch.uint8, device='cuda')
            _lib.check(lib.b2fwi_w1d_misfit(_ptr(fd), _ptr(gd), None, nt, nrec, 1, ctypes.c_double(float(self.gamma)),
                                            _ptr(adj), _ptr(fval), _ptr(scratch), _stream()))
        return float(fval.item()), adj.cpu().numpy().astype(np.float64)


class Misfit(object):
    """misfit/misfit.py:106-111."""

    def __init__(self, operator):
        self.operator = operator

    def __call__(self, x, y):
        return self.operator(x, y)
