"""oracle/ -- TEST INFRASTRUCTURE ONLY (CPU checkers for the sm_100a kernels).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product package (opensplat_b200) never does.
"""
