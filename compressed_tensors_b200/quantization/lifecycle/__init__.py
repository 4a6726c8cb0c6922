from .forward import *  # noqa: F401,F403
from .initialize import *  # noqa: F401,F403
from .apply import *  # noqa: F401,F403
from .compressed import *  # noqa: F401,F403
from .helpers import *  # noqa: F401,F403
