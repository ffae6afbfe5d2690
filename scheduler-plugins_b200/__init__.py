"""b200sched — B200-native batched Filter/Score engine behind the scheduler-plugins interface.

Layout
  csrc/    hand-written sm_100a kernels + the C-ABI (include/b200sched.h) -> lib/libb200sched.so
  host/    C++ mirror of the reference's plugin interface (flattening + per-cycle lookup)
  engine.py  ctypes binding over the C-ABI used by the tests and the bench harness
  synth.py   seeded synthetic snapshot generator (one generator, three consumers)

The CUDA library is mandatory: importing `engine` without it raises, there is no CPU fallback.
"""
__all__ = ["engine", "synth"]
