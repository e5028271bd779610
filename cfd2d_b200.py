"""Import shim: the package directory is named ``cfd-2d_b200`` (after the reference, zhrv/cfd-2d),
which is not a valid Python identifier.  ``import cfd2d_b200`` loads that directory as a package."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "cfd-2d_b200")]
__file__ = _os.path.join(__path__[0], "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
