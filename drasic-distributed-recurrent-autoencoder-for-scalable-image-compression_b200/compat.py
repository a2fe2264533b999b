"""torch-2.x compatibility for the reference's caller scripts.  The reference was written for
PyTorch 1.0 (README.md:12); on the installed torch 2.11 its train_model.py:179-182 fails because
`ReduceLROnPlateau(verbose=True)` is no longer accepted.  config.init() applies this shim, so a caller that
does `import config; config.init()` first -- as train_model.py:1-3 and test_model.py do -- runs unchanged.
(`utils.load` passes weights_only=False for the same reason: the reference checkpoints pickle config dicts.)"""
import inspect

import torch.optim.lr_scheduler as _sched

_applied = False


def apply():
    global _applied
    if _applied:
        return
    _applied = True
    cls = _sched.ReduceLROnPlateau
    if 'verbose' in inspect.signature(cls.__init__).parameters:
        return                                   # an older torch that still has the argument
    orig_init = cls.__init__

    def __init__(self, optimizer, *args, verbose=False, **kwargs):
        orig_init(self, optimizer, *args, **kwargs)
        self.verbose = verbose

    cls.__init__ = __init__
