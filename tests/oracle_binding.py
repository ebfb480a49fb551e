"""Re-export of the oracle binding for the tests (the oracle is test infrastructure, see oracle/binding.py)."""
from oracle.binding import *  # noqa: F401,F403
from oracle.binding import lib, build  # noqa: F401
