"""Import shim: the product package lives in ``asteroid-spring-sims_b200/`` (hyphenated, as the layout contract names
it); Python cannot import that spelling, so this package points its search path there and runs its __init__."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "asteroid-spring-sims_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
