from said_b200.util.audio import FittedWaveform, fit_audio_unet, load_audio  # noqa: F401
