"""Import alias: the package lives in the directory `index-tts_b200/` (the name the build
contract asks for), which is not a valid Python identifier.  `import indextts_b200` resolves
to that directory."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "index-tts_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f
