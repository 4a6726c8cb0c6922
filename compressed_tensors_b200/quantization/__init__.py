from .quant_args import *  # noqa: F401,F403
from .quant_scheme import *  # noqa: F401,F403
from .quant_config import *  # noqa: F401,F403
from .quant_metadata import *  # noqa: F401,F403
from .utils import *  # noqa: F401,F403
from .lifecycle import *  # noqa: F401,F403
