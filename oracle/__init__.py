"""CPU oracle — TEST INFRASTRUCTURE ONLY (see oracle/rbd_oracle.hpp).

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs; never from pinocchio_b200/.
"""
from .oracle import Oracle, build_oracle  # noqa: F401
