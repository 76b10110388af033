import sys; sys.path.insert(0, '.')
from pour_over_coffee_lbm_b200.engine import D3Q19Engine
e = D3Q19Engine(8, 8, 8, compat="physical")
print("selftest", e.selftest_math())
