"""athena_b200 -- B200-native engine for Atmospheric Athena's per-timestep hot path.

The product is the C-ABI library `csrc/libathena_b200.so` (hand-written FP64 CUDA for sm_100a, declared in
include/athena_b200.h).  This Python package is the host-side mirror of the reference's interface for that
path (main.c loop, par.c lookups, problem-file hooks) used by tests and bench.py; it does no arithmetic on
the state itself.
"""
