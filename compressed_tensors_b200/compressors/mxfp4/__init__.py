from .base import *  # noqa: F401,F403
