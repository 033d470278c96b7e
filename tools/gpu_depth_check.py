import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.gpu_check import parity
r = {}
r["depth_small"] = parity("depth_small", 3000, 72, 104, 3, depth=True)
r["depth_sh"] = parity("depth_sh", 2000, 64, 80, 3, depth=True, use_sh=True, deg=3, M=16)
print(r)
