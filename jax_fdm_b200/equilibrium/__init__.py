from .fdm import *  # noqa: F401,F403
from .models import *  # noqa: F401,F403
from .solvers import *  # noqa: F401,F403
from .sparse import *  # noqa: F401,F403
from .states import *  # noqa: F401,F403
from .structures import *  # noqa: F401,F403
from . import ops  # noqa: F401
