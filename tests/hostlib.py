"""Test-side alias of the package's host-library binding."""
from helpers import pkg

_m = pkg().hostapi
globals().update({k: getattr(_m, k) for k in dir(_m) if not k.startswith("__")})
