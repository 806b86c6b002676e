"""Import shim: the product package lives in ``climaatmos.jl_b200/`` (a directory name Python
cannot import directly because of the dot); this module re-exports it as ``climaatmos_jl_b200``."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "climaatmos.jl_b200")]
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
