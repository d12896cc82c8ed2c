"""CPU oracle for the Kinetic-GAN ST-GCN hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
import this package, and only as the checker / the timed CPU baseline - never as part of the
product path (kinetic-gan_b200/), which fails loudly when its CUDA library is missing.

Parity pin: the reference holds no tests or golden vectors, so the oracle is
pinned against outputs of the UNMODIFIED reference imported in the build container
(oracle/ref_shim.py) - see tests/golden/make_golden.py and tests/test_oracle_golden.py.
"""
