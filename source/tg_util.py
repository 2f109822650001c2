"""The reference's source/tg_util.py with RC() bound to the str.translate form (same result, same KeyError; tg_util.py:433-434)."""
from source._overlay import exec_reference

exec_reference(__name__, globals())

from telogator2_b200.tg_util import RC  # noqa: E402,F401
