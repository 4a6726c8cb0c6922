from .base import *  # noqa: F401,F403
from .dense import *  # noqa: F401,F403
from .naive_quantized import *  # noqa: F401,F403
from .pack_quantized import *  # noqa: F401,F403
from .sparse import *  # noqa: F401,F403
from .format import *  # noqa: F401,F403
from .model_compressors import *  # noqa: F401,F403
