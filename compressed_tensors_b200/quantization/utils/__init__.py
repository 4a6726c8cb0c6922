from .helpers import *  # noqa: F401,F403
