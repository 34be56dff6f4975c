"""Minimal `diffusers` stand-in, used only when the real package is not installed: the reference's
scripts import `DDIMScheduler` from it to pass as `noise_scheduler=` (script/inference.py:5, 153)."""
from said_b200.scheduler import DDIMScheduler, DDPMScheduler, SchedulerMixin  # noqa: F401
