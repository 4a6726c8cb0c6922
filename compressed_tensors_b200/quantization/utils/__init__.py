from .helpers import *  # noqa: F401,F403
from .mxfp_utils import *  # noqa: F401,F403
