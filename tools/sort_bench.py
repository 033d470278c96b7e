#!/usr/bin/env python
"""Developer diagnostic: the library's radix sort (sagars_sort_pairs) vs cub::DeviceRadixSort vs torch.sort --
correctness (stability included) on awkward sizes, then CUDA-event timings at the bench's instance counts."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from seganygaussians_b200 import _lib  # noqa: E402


def run(lib, keys, vals, bits, use_cub, temp):
    ko, vo = torch.empty_like(keys), torch.empty_like(vals)
    rc = lib.sagars_sort_pairs(0, keys.numel(), bits, keys.data_ptr(), vals.data_ptr(), ko.data_ptr(), vo.data_ptr(),
                               temp.data_ptr(), use_cub, torch.cuda.current_stream().cuda_stream)
    assert rc == 0, _lib.last_error()
    return ko, vo


def main():
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    ok_all = True
    for n, bits in [(1, 41), (255, 41), (4096, 45), (4097, 46), (100003, 45), (3000000, 45), (1000000, 33), (50000, 64)]:
        g = torch.Generator().manual_seed(n)
        hi = torch.randint(0, 1 << min(bits - 32, 13), (n,), generator=g, dtype=torch.int64) if bits > 32 else torch.zeros(n, dtype=torch.int64)
        lo = torch.randint(0, 1 << 20, (n,), generator=g, dtype=torch.int64)   # many ties -> stability matters
        keys = ((hi << 32) | lo).to(dev)
        vals = torch.arange(n, dtype=torch.int32, device=dev)
        temp = torch.empty(int(lib.sagars_sort_temp_bytes(n)), dtype=torch.uint8, device=dev)
        ko, vo = run(lib, keys, vals, bits, 0, temp)
        torch.cuda.synchronize()
        sk, idx = torch.sort(keys, stable=True)
        ok = bool(torch.equal(ko, sk) and torch.equal(vo, vals[idx]))
        ok_all &= ok
        print(f"  n={n:8d} bits={bits}: own==torch.sort(stable) {ok}")
    for n in (2_856_836, 16_366_754):
        g = torch.Generator().manual_seed(1)
        keys = ((torch.randint(0, 8160, (n,), generator=g, dtype=torch.int64) << 32) |
                torch.randint(0x3E000000, 0x42000000, (n,), generator=g, dtype=torch.int64)).to(dev)
        vals = torch.arange(n, dtype=torch.int32, device=dev)
        temp = torch.empty(int(lib.sagars_sort_temp_bytes(n)), dtype=torch.uint8, device=dev)
        for name, use_cub in (("own", 0), ("cub", 1)):
            for _ in range(3):
                run(lib, keys, vals, 45, use_cub, temp)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                run(lib, keys, vals, 45, use_cub, temp)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            print(f"  n={n:9d} 45 bits {name}: {ms:.3f} ms  ({n / ms / 1e6:.1f} G pairs/s; includes the input copy of the stand-alone entry)")
    print("OK" if ok_all else "FAIL")
    return 0 if ok_all else 1


if __name__ == "__main__":
    sys.exit(main())
