"""Import alias: `import ssr_b200` loads the package that lives in `stuttering-speech-representation_b200/`
(a directory name with hyphens cannot be imported directly)."""
import os as _os

_pkg_dir = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         "stuttering-speech-representation_b200")
__path__ = [_pkg_dir]
with open(_os.path.join(_pkg_dir, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_pkg_dir, "__init__.py"), "exec"))
