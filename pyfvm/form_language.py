"""``from pyfvm.form_language import *`` [README:32, 88]."""
from pyfvm_b200.form_language import *  # noqa: F401,F403
from pyfvm_b200.form_language import __all__  # noqa: F401
