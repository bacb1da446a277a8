"""Import shim: `kerneldensityestimate.jl_b200/` is not a valid Python identifier, so this tiny
package extends its search path with that directory; `import kde_b200` then resolves every
submodule (api, _lib, dist, build) from kerneldensityestimate.jl_b200/."""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                                 "kerneldensityestimate.jl_b200"))

from .api import *  # noqa: E402,F401,F403
from .api import __all__  # noqa: E402,F401
