from .fgfuncs import *  # noqa
from .imgprocess import *  # noqa
from .maskprocess import *  # noqa
from .visualize import *  # noqa
from .temporal import *  # noqa
from .fileio import *  # noqa
from .region_fill import *  # noqa
