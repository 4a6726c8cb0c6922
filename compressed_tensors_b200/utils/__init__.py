from .helpers import *  # noqa: F401,F403
from .module import *  # noqa: F401,F403
from .impl_backend import ImplBackend  # noqa: F401
from .match import *  # noqa: F401,F403
from .internal import InternalModule  # noqa: F401
