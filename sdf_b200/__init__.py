"""sdf_b200 -- B200-native mesh generation behind fogleman/sdf's Python API.

Star-exports the same public names as the reference package (sdf/__init__.py:1-27):
3-D names win over 2-D names of the same spelling, as there.
"""
from . import d2, d3, ease

from .util import *            # noqa: F401,F403

from .d2 import *              # noqa: F401,F403

from .d3 import *              # noqa: F401,F403

from .texture import (         # noqa: F401  (text / image are host nodes; Mesh needs pyopenvdb and raises)
    Mesh,
    image,
    measure_image,
    measure_text,
    text,
)

from .core import (            # noqa: F401
    generate,
    save,
    sample_slice,
    show_slice,
)

from .stl import (             # noqa: F401
    write_binary_stl,
)

from .ir import NotLowerableError   # noqa: F401
