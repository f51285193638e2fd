"""oracle -- CPU restatement of the reference's hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package, and only as the checker.  The product (libnct.so and the
neural-color-transfer_b200 package) never imports it.

Parts:
  pm_oracle.c      C: XORWOW, NNF init/upsample, L2 norm, deterministic PatchMatch
                   (+ the reference-order distance and a reference-semantics serial PatchMatch)
  pm.py            ctypes wrappers over liboracle_pm.so
  synth.py         the synthetic inputs of SURVEY.md section 8(d)
"""
from .pm import *  # noqa: F401,F403
