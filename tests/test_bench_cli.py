"""bench.py's developer switches must name things the library knows: every --fwd-kernel / --bwd-kernel / --binning choice is accepted
by the rasterizer's own setters, and the documented contract keys of the JSON line are spelled the way bench.py emits them.  CPU only
(nothing is launched)."""
import argparse
import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _choices(bench, flag, monkeypatch):
    seen = {}
    orig = argparse.ArgumentParser.add_argument

    def spy(self, *names, **kw):
        if flag in names:
            seen["choices"] = list(kw.get("choices") or [])
            seen["default"] = kw.get("default")
        return orig(self, *names, **kw)

    monkeypatch.setattr(argparse.ArgumentParser, "add_argument", spy)
    monkeypatch.setattr(sys, "argv", ["bench.py"])
    bench.parse()
    return seen


def test_kernel_switches_are_known_to_the_rasterizer(bench, monkeypatch):
    from seganygaussians_b200 import rasterizer as R
    try:
        for f in _choices(bench, "--fwd-kernel", monkeypatch)["choices"]:
            R.set_blend_kernels(forward=f)
        for b in _choices(bench, "--bwd-kernel", monkeypatch)["choices"]:
            R.set_blend_kernels(backward=b)
        for m in _choices(bench, "--binning", monkeypatch)["choices"]:
            if m != "default":
                R.set_binning(m)
    finally:
        R.set_blend_kernels()
        R.set_binning()


def test_defaults_follow_the_contract(bench, monkeypatch):
    monkeypatch.setattr(sys, "argv", ["bench.py"])
    a = bench.parse()
    assert a.gpus == 1 and a.warmup >= 3 and a.steps >= 200 and a.impl == "ours" and a.workload == "c2"
    wl = bench.WORKLOADS["c2"]
    assert (wl["P"], wl["H"], wl["W"], wl["K"]) == (1_000_000, 1080, 1920, 32)          # BASELINE.json configs[1]
    c4 = bench.WORKLOADS["c4"]
    assert c4["P"] == 3_000_000 and c4["K"] == 32                                       # BASELINE.json configs[3]
    names = bench.kernel_names(a, wl["K"])
    assert set(names) == {"forward", "backward", "binning"}
