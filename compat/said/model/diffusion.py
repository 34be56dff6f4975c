from said_b200.model.diffusion import *  # noqa: F401,F403
from said_b200.model.diffusion import SAID, SAID_UNet1D, SAIDInferenceOutput, SAIDNoiseAdditionOutput  # noqa: F401
