"""oracle/ -- TEST INFRASTRUCTURE ONLY (CPU checkers for the CUDA path).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  See oracle/oracle.h for the parity status of each function.
"""
