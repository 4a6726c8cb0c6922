from .model_compressor import *  # noqa: F401,F403
