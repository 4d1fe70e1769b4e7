"""CPU fp32 oracle of the FMC denoising hot path -- TEST INFRASTRUCTURE, never the product path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import this
package.  PARITY UNPINNED: the reference holds no golden vectors and cannot be imported here (it needs
diffusers==0.24.0), see oracle/diffusers_restated.py and DESIGN.md.
"""
