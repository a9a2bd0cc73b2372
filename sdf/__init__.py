"""Drop-in alias: ``from sdf import *`` resolves to the B200-native package."""
from sdf_b200 import *          # noqa: F401,F403
from sdf_b200 import d2, d3, core, dn, ease, stl, util, progress   # noqa: F401
