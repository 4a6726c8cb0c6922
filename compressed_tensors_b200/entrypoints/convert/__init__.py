from .converters import *  # noqa: F401,F403
from .convert_file import *  # noqa: F401,F403
from .convert_checkpoint import *  # noqa: F401,F403
