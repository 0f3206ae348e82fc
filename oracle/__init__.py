"""CPU oracle for the BPE-encode hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this package.  See bpe_oracle.c for what is restated and oracle/README.md for how it
is pinned against the reference engine.
"""
from .oracle import Oracle, PAT_R50K, PAT_CL100K, PAT_O200K, pattern_id, build  # noqa: F401
