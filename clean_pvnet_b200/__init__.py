"""Import alias: `import clean_pvnet_b200` loads the package that lives in ../clean-pvnet_b200/
(a hyphen is not importable).  No code lives here."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "clean-pvnet_b200")]
with open(_os.path.join(__path__[0], "__init__.py")) as _fh:
    exec(compile(_fh.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
del _os, _fh
