"""oracle - CPU checkers for the CUDA path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package.
The product (pagmo2_b200/, include/) never does.
"""
