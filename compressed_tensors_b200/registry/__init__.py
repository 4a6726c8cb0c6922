from .registry import *  # noqa: F401,F403
