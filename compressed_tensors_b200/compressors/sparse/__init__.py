from .bitmask import *  # noqa: F401,F403
