"""`diffusers` stand-in for machines without the real package: the reference's scripts import `DDIMScheduler` from it to
pass as `noise_scheduler=` (script/inference.py:5, 153).

`compat/` has to come first on the path for the `said` / `dataset` shims, which would also put this stand-in ahead of an installed
`diffusers`.  So it looks for another `diffusers` further down `sys.path` and, if there is one, hands the import over to it (the
real package is what ends up in `sys.modules["diffusers"]`); only otherwise does it export the restated schedulers."""
import importlib.machinery
import importlib.util
import os
import sys

_here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_others = [p for p in sys.path if os.path.abspath(p or ".") != _here]
_spec = importlib.machinery.PathFinder.find_spec("diffusers", _others)
if _spec is not None and _spec.loader is not None:
    _real = importlib.util.module_from_spec(_spec)
    sys.modules["diffusers"] = _real            # the import statement that triggered us returns what sys.modules holds
    _spec.loader.exec_module(_real)
else:
    from said_b200.scheduler import DDIMScheduler, DDPMScheduler, SchedulerMixin  # noqa: F401
