"""CPU oracle (test infrastructure only) -- see swiftly_oracle.py."""
from .swiftly_oracle import *  # noqa: F401,F403
from .swiftly_oracle import OracleCore  # noqa: F401
