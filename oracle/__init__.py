"""TEST INFRASTRUCTURE ONLY: CPU oracle for tt_irt1 (see oracle/tt_irt1_oracle.c).

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Nothing under tt-irt_b200/ imports this package.
"""
from . import parity  # noqa: F401
from . import samplers_oracle  # noqa: F401
from .oracle import (  # noqa: F401
    build, have_ref, oracle_run, oracle_sweep, ref_run, ref_lib_path, ORACLE_DIR,
)
