from .assign import *  # noqa: F401,F403
from .utils import *  # noqa: F401,F403
from .module_parallel import *  # noqa: F401,F403
