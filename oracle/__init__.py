"""CPU oracle for the SmoothParticleNets hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  The product (smoothparticlenets_b200) never does; it fails loudly without its CUDA
library instead of falling back to anything here.
"""
