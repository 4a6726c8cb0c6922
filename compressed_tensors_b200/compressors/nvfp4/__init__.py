from .base import *  # noqa: F401,F403
from .helpers import *  # noqa: F401,F403
