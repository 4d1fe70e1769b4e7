"""CPU fp32 oracle of the FMC denoising hot path -- TEST INFRASTRUCTURE, never the product path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import this
package.

Pinning status: the reference holds no golden vectors or tests of its own and cannot be imported as-is (it needs
diffusers==0.24.0, not installable offline).  The restatement of everything under fmc/ IS pinned: tests/golden/
make_golden.py executes the reference's own fmc/* sources in the build container (third-party layer shimmed) and
tests/test_golden.py holds this oracle to those outputs at 2e-5.  PARITY UNPINNED only at the diffusers boundary:
oracle/diffusers_restated.py restates diffusers 0.24.0 classes from their published behaviour
and is cross-checked against torch primitives (tests/test_oracle.py), not against diffusers itself.
"""
