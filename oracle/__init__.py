"""CPU oracle for the D3Q19 hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  The product (pour_over_coffee_lbm_b200) never does.
"""
