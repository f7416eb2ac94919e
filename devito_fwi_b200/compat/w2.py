"""Stand-in for the ``w2`` extension module that misfit/bfm.py:1 imports unconditionally but the reference does
not ship. Only the pure-Python ``bfm`` class needs it; the L2 and 1-D Wasserstein misfits do not."""


class BFM(object):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("the `w2` back-and-forth extension is not part of the reference tree; "
                                  "use misfit.least_square or qWasserstein(method='1d')")
