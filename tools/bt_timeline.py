"""Developer tool: phase timeline of the tcgen05 backward (library built with EXTRA_NVCCFLAGS=-DSAGARS_BT_TIMELINE into
seganygaussians_b200/lib/timeline/libsagars.so; SAGARS_LIBRARY selects it).  Prints, per role, the share of its time per phase."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tests import common
from seganygaussians_b200 import synthetic, rasterizer as R, _lib
lib = _lib.load()
P, H, W, K = 1_000_000, 1080, 1920, 32
sc = synthetic.scene(P, H, W, K)
R.set_blend_kernels(backward="tc")
common.run_torch_impl("ours", sc, K)
buf = (ctypes.c_ulonglong * 32)()
lib.sagars_bt_timeline_read(buf)            # clear (warm-up included the counters)
common.run_torch_impl("ours", sc, K)
torch.cuda.synchronize()
assert lib.sagars_bt_timeline_read(buf) == 0
v = list(buf)
names = {0: "setup (all warps, measured on consumer warp 0)", 1: "consumer: -> prologue", 2: "consumer: prologue (F tiles of batches 0, 1)",
         3: "consumer: table wait + F load issue", 4: "consumer: wait S", 5: "consumer: ld S + pass 1", 6: "consumer: wait B3 free",
         7: "consumer: pass 2 + B3 stores", 16: "consumer: F rows -> tile (waits for the load issued at the top of the batch)", 17: "consumer: fence.proxy.async", 8: "consumer: arrive", 9: "consumer: wait for CTA end",
         10: "issuer/epilogue warp 4: start", 11: "warp 4: wait B3 full", 12: "warp 4: wait D3 free", 13: "warp 4: issue 8 MMAs",
         14: "warp 4: wait product", 15: "warp 4: rows -> global", 20: "producer: start", 21: "producer: selection / publish",
         22: "producer: S issue (8 MMAs)", 23: "producer: idle"}
for grp, title in (((0, 1, 2, 3, 4, 5, 6, 7, 16, 17, 8, 9), "consumer warp 0"), ((10, 11, 12, 13, 14, 15), "warp 4"), ((20, 21, 22, 23), "producer warp 7")):
    tot = sum(v[i] for i in grp) or 1
    print(f"== {title}: {tot / 1e6:.1f} M cycles summed over CTAs")
    for i in grp:
        print(f"   {100 * v[i] / tot:5.1f} %  {v[i] / 1e6:9.1f} M  {names[i]}")
