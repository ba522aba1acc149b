"""TEST INFRASTRUCTURE ONLY: ctypes loader for the CPU oracle (oracle/liblaser_oracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  Nothing under laser_b200/ does.
"""
from .oracle import *  # noqa: F401,F403
