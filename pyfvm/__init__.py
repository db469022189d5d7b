"""``import pyfvm`` shim: the reference's package name bound to the B200 engine,
so README code runs unchanged (``import pyfvm``; ``from pyfvm.form_language
import ...``)."""
from pyfvm_b200 import *  # noqa: F401,F403
from pyfvm_b200 import __version__, form_language, fvm_matrix  # noqa: F401
