"""CPU oracle of the fredholm rendering core -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  The product (fredholm_b200/) never does.
"""
