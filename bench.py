#!/usr/bin/env python
"""bench.py -- fwd+bwd throughput of the Gaussian feature rasterizer on BASELINE.json's headline config.

    python bench.py --gpus N --steps K --warmup W [--impl ours|reference]

* workload (config.workload): BASELINE.json configs[1] = synthetic 1M Gaussians, 1080x1920, K=32, one
  camera per GPU per step (SYN(P,H,W,K,cam) of BASELINE.md section 2.2; camera index = rank).
* a "step" = one forward + backward of the rasterizer over one camera on every rank, plus (N>1) ONE NCCL
  all-reduce of the per-Gaussian feature gradient dL_dcolors [P,K].  Weak scaling: per-GPU work is fixed.
* metric = Gaussians*pixels/s = (sum over ranks of P*H*W) * K_steps / time; time = CUDA events on the launching
  stream, barrier + synchronize on both sides, max over ranks.
* `value`  : Gaussian parameters, upstream gradient and camera resident in HBM, operator called directly.
* `e2e`    : the same step through the reference-shaped public API as a user drives it -- per step the camera
             (view / projection / centre / background: the only host-side inputs the reference API has; the
             Gaussian parameters are resident model state exactly as in train_contrastive_feature.py) is copied
             from pinned host memory, the loss (sum(image * dL)) is read back to the host.
* `roofline`: HBM roofline of the dominant kernel, from per-stage CUDA-event timings taken by the library on its
             launch stream inside the timed region, with algorithmic bytes per launch as defined in DESIGN.md.
* `cpu_baseline`: the CPU oracle port (oracle/sagars_oracle.c, OpenMP over tiles) timed on the host cores on the
             same workload (rank 0, N=1 only).
* `--impl reference`: the UNMODIFIED reference extension (oracle/_ref, rebuilt for sm_100a from /root/reference by
             oracle/build_ref.py) through its own GaussianRasterizer API on the same inputs.  NOTE: the reference has
             no CPU implementation of this path -- its implementation IS the CUDA extension, and the north star's
             ">= 3x the reference CUDA rasterizer" is a ratio against exactly this arm.  If oracle/_ref is absent the
             arm falls back to the CPU oracle port and says so.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

WORKLOADS = {
    # BASELINE.json configs[1]
    "c2": dict(P=1_000_000, H=1080, W=1920, K=32, desc="synthetic 1M Gaussians, 1080x1920, K=32 affinity features, 1 camera/GPU, fwd+bwd"),
    # BASELINE.json configs[2]-like (parity-test case, not the bench line): 5M Gaussians
    "c3": dict(P=5_000_000, H=1036, W=1600, K=32, desc="synthetic 5M Gaussians, 1036x1600, K=32 (garden-like), 1 camera/GPU, fwd+bwd"),
    # configs[1] at K=3 (the BASE variant's channel count; developer A/B of the forward kernels, never a bench line)
    "c2_k3": dict(P=1_000_000, H=1080, W=1920, K=3, desc="synthetic 1M Gaussians, 1080x1920, K=3, 1 camera/GPU, fwd+bwd"),
    # small variant for quick local checks (never a bench line)
    "tiny": dict(P=20_000, H=270, W=480, K=32, desc="tiny smoke workload"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--fwd-kernel", default="default", choices=["default", "tile", "warp_any"], help="developer A/B switch (rasterizer.set_blend_kernels)")
    ap.add_argument("--bwd-kernel", default="default", choices=["default", "tile", "tc"], help="developer A/B switch")
    ap.add_argument("--overlap-allreduce", action="store_true",
                    help="N > 1, developer switch: all-reduce on a side stream, gating only the next forward's blend stage "
                         "(its geometry stages overlap the exchange); default: the all-reduce serialises with the step")
    ap.add_argument("--allreduce", default="nccl", choices=["nccl", "multimem"],
                    help="N > 1, developer switch: 'multimem' = the library's own all-reduce over the NVSwitch multicast mapping")
    ap.add_argument("--binning", default="radix", choices=["radix", "tile_sort"], help="developer A/B switch (rasterizer.set_binning)")
    return ap.parse_args()


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region, read through NVML (the quantities
    `nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.*` prints; B200_PROFILING.md recipe).
    Samples are taken explicitly from the launching thread while the GPU is busy with the timed steps (after
    the launches of the middle step and of the last step, before the closing synchronise).  A polling child /
    thread is deliberately NOT used: `nvidia-smi -lms 50` was measured to slow this launch-heavy step 3x and a
    100 ms NVML thread still added up to ~1 ms/step of jitter, which would falsify the number it guards."""

    def __init__(self, index: int):
        self.index, self.rows, self.ok = index, [], False
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")   # CUDA ordinal -> NVML index
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].strip().isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:   # pragma: no cover
            self.err = repr(e)

    def sample(self):
        if not self.ok:
            return
        nv = self.nv
        try:
            sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
            rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            self.rows.append((sm, int(rs)))
        except Exception:
            pass

    def result(self):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "err", "?")]}
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        reasons = sorted(n for n, bit in names.items() if any(r[1] & bit for r in self.rows))
        sm = [r[0] for r in self.rows]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(self.max_sm), "reasons": reasons,
                "samples": len(sm), "source": "nvml, sampled while the timed steps execute"}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def algorithmic_bytes(P, R, HW, T, C):
    """DESIGN.md section 'algorithmic bytes' (SURVEY.md Appendix C): each datum crosses HBM once."""
    per_stage = {
        "preprocess": P * (44 + 60),
        "scan_block_sums": P * 8,
        "duplicate_keys": P * 20 + R * 12,
        "radix_sort": R * 24,
        "tile_ranges": R * 8 + T * 8,
        "render_forward": R * (28 + 4 * C) + HW * (4 * C + 8),
        "render_backward": R * (28 + 4 * C) + HW * (4 * C + 8) + P * (4 * C + 24),
        "geom_backward": P * (96 + 64),
    }
    return per_stage, sum(per_stage.values())


def main():
    a = parse()
    wl = WORKLOADS[a.workload]
    P, H, W, K = wl["P"], wl["H"], wl["W"], wl["K"]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # Reference arm under torchrun: the reference's rasterizer is a CUDA extension, so it gets the same treatment as ours
    # (one camera per rank + the all-reduce of dL_dcolors, i.e. what train_contrastive_feature.py would do under DDP).
    # Only the CPU-port fallback (no oracle/_ref on this box) runs on rank 0 alone.
    ref_is_cuda = False
    if a.impl == "reference":
        from tests import common as _c
        ref_is_cuda = _c.have_ref("cf") and K == 32
        if world > 1 and rank != 0 and not ref_is_cuda:
            return 0
    import torch.distributed as dist
    use_dist = world > 1 and (a.impl == "ours" or ref_is_cuda)
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if use_dist:
        dist.init_process_group("nccl", device_id=dev)

    from seganygaussians_b200 import synthetic
    cam_index = rank % 8
    sc = synthetic.scene(P, H, W, K, cam=cam_index)
    g = sc.gauss

    # ---------------- implementation under test ----------------
    ref_kind = None
    if a.impl == "ours":
        from seganygaussians_b200 import rasterizer as R, _lib
        _lib.load()
        R.set_blend_kernels(forward=a.fwd_kernel, backward=a.bwd_kernel)
        R.set_binning(a.binning)
        Settings, Rast = R.GaussianRasterizationSettings, R.GaussianRasterizerContrastiveF
    else:
        from tests import common
        if common.have_ref("cf") and K == 32:
            mod = common.ref_module("cf")
            Settings, Rast = mod.GaussianRasterizationSettings, mod.GaussianRasterizer
            ref_kind = "reference-cuda-ext"
        else:
            return reference_cpu_arm(a, wl, sc)

    means3D = g.means3D.to(dev).requires_grad_(True)
    means2D = torch.zeros(P, 3, device=dev, requires_grad=True)
    opac = g.opacities.to(dev).requires_grad_(True)
    scales = g.scales.to(dev).requires_grad_(True)
    rots = g.rotations.to(dev).requires_grad_(True)
    colors = g.colors.to(dev).requires_grad_(True)
    leaves = (means3D, means2D, opac, scales, rots, colors)
    dL = sc.dL_dout.to(dev)
    c = sc.cam
    bg_d = torch.zeros(K, device=dev)
    view_d, proj_d, campos_d = c.world_view_transform.to(dev), c.full_proj_transform.to(dev), c.camera_center.to(dev)
    # pinned host copies of the per-step inputs (e2e leg)
    view_h, proj_h, campos_h, bg_h = (t.clone().pin_memory() for t in (c.world_view_transform, c.full_proj_transform,
                                                                        c.camera_center, torch.zeros(K)))
    h2d_bytes = sum(t.numel() * 4 for t in (view_h, proj_h, campos_h, bg_h))

    def settings(view, proj, campos, bg):
        return Settings(image_height=H, image_width=W, tanfovx=c.tanfovx, tanfovy=c.tanfovy, bg=bg, scale_modifier=1.0,
                        viewmatrix=view, projmatrix=proj, sh_degree=0, campos=campos, prefiltered=False, debug=False)

    overlap = bool(getattr(a, "overlap_allreduce", False)) and a.impl == "ours"
    reducer = None
    if overlap and world > 1:
        from seganygaussians_b200.data_parallel import FeatureGradReducer
        reducer = FeatureGradReducer(side_stream=True)
    own_allreduce = None
    if a.impl == "ours" and use_dist and a.allreduce == "multimem":
        from seganygaussians_b200.data_parallel import MulticastAllReduce
        own_allreduce = MulticastAllReduce(P * K, dev)

    def reduce_grad(t):
        if own_allreduce is not None:
            own_allreduce.all_reduce_(t)
        else:
            dist.all_reduce(t)

    rs_resident = settings(view_d, proj_d, campos_d, bg_d)
    rast_resident = Rast(raster_settings=rs_resident)
    last = {}

    def step_resident():
        for t in leaves:
            t.grad = None
        color, radii = rast_resident(means3D=means3D, means2D=means2D, opacities=opac, shs=None, colors_precomp=colors,
                                     scales=scales, rotations=rots, cov3D_precomp=None)
        last["num_rendered"] = int(color.grad_fn.num_rendered)
        last["radii"] = radii
        color.backward(dL)
        if use_dist:
            if reducer is not None:
                reducer.wait()                               # at most one exchange in flight
                reducer.reduce_async(colors.grad)            # side stream, behind everything queued so far
                R.set_blend_wait_event(reducer.ready_event())
            else:
                reduce_grad(colors.grad)

    def step_e2e():
        for t in leaves:
            t.grad = None
        view = view_h.to(dev, non_blocking=True); proj = proj_h.to(dev, non_blocking=True)
        campos = campos_h.to(dev, non_blocking=True); bg = bg_h.to(dev, non_blocking=True)
        rast = Rast(raster_settings=settings(view, proj, campos, bg))
        color, radii = rast(means3D=means3D, means2D=means2D, opacities=opac, shs=None, colors_precomp=colors,
                            scales=scales, rotations=rots, cov3D_precomp=None)
        loss = (color * dL).sum()
        loss.backward()
        if use_dist:
            reduce_grad(colors.grad)
        return float(loss.item())   # device -> host read of the step's result

    def barrier():
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps, sampler=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn()
            if sampler is not None and (i == steps // 2 or i == steps - 1):
                sampler.sample()          # the GPU is still executing this step's backward
        if reducer is not None:
            reducer.wait()                # the last exchange belongs to the timed region
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if use_dist:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---------------- warm-up, then the timed regions ----------------
    sampler = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(max(a.warmup, 3)):
        step_resident()
    step_e2e()
    barrier()
    if a.impl == "ours":
        _lib.reset_launch_count()
        _lib.profile_read(reset=True)
        _lib.profile_enable(True)
    ms_total = timed(step_resident, a.steps, sampler)
    stage = None
    launches = None
    if a.impl == "ours":
        _lib.profile_enable(False)
        stage = _lib.profile_read(reset=True)
        launches = _lib.launch_count()
    ms_e2e = timed(step_e2e, a.steps)
    clocks = sampler.result() if sampler else None

    n_used = world if use_dist else 1
    gp_per_step = float(P) * H * W * n_used
    value = gp_per_step * a.steps / (ms_total * 1e-3)
    e2e_value = gp_per_step * a.steps / (ms_e2e * 1e-3)

    if rank != 0:
        if use_dist:
            dist.destroy_process_group()
        return 0

    # ---------------- work counters of rank 0's camera ----------------
    R_inst = last["num_rendered"]
    radii_vis = int((last["radii"] > 0).sum().item())
    T_tiles = ((W + 15) // 16) * ((H + 15) // 16)
    # work-proportional counters: S = pair tests a tile-per-CTA traversal would make at most
    # (256 threads x every instance of the tile); n_contrib summed = pairs up to each pixel's last contributor.
    S_pairs = 256 * R_inst
    pairs_to_last = None
    if a.impl == "ours":
        try:   # one extra UNTIMED forward: n_contrib lives in the call's opaque image-state buffer
            color_x, _ = rast_resident(means3D=means3D, means2D=means2D, opacities=opac, shs=None, colors_precomp=colors,
                                       scales=scales, rotations=rots, cov3D_precomp=None)
            img_state = color_x.grad_fn.saved_tensors[-1]
            il = _lib.image_layout(W, H)
            raw = img_state[il.n_contrib: il.n_contrib + 4 * H * W]
            pairs_to_last = int(raw.view(torch.int32).sum(dtype=torch.int64).item())
        except Exception:
            pairs_to_last = None
    line = {
        "metric": f"fwd+bwd Gaussians*pixels/s @K={K}", "value": value, "unit": "Gaussian*pixel/s",
        "n_gpus": world if (use_dist or a.impl != "ours") else a.gpus, "steps": a.steps, "warmup": max(a.warmup, 3),
        "ms_per_step": ms_total / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "impl": a.impl,
        "config": {"workload": wl["desc"], "P": P, "H": H, "W": W, "K": K, "cameras_per_step": n_used,
                   "parallelism": f"camera-dp{n_used}" + ("+allreduce(dL_dcolors)" if use_dist else "") +
                                  ("+overlapped with the next forward's geometry stages" if reducer is not None else "") +
                                  (" [own multimem all-reduce]" if own_allreduce is not None else ""),
                   "l2": "inputs_exceed_l2 (features 128 MB + upstream gradient 265 MB + image 265 MB >> 126 MB L2)",
                   "P_visible": radii_vis, "R_instances": R_inst,
                   "kernels": {"forward": {"default": "mma.sync warp kernel at K=32, fp32 tile kernel otherwise", "tile": "tcgen05 tile kernel",
                                           "warp_any": "mma.sync warp kernel for every K"}[a.fwd_kernel],
                               "backward": {"default": "mma.sync warp kernel", "tile": "mma.sync tile kernel", "tc": "tcgen05 / TMEM pixel-group kernel"}[a.bwd_kernel],
                               "binning": a.binning},
                   "S_pair_tests_upper_bound": S_pairs, "pairs_up_to_last_contributor": pairs_to_last},
        "e2e": {"value": e2e_value, "unit": "Gaussian*pixel/s", "ms_per_step": ms_e2e / a.steps,
                "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 8,
                "note": "camera from pinned host memory each step; loss scalar + num_rendered read back"},
        "clocks": clocks,
    }
    if a.impl == "ours":
        peak, peak_src = load_peaks()
        per_stage_bytes, step_bytes = algorithmic_bytes(P, R_inst, H * W, T_tiles, K)
        dom = max(stage.items(), key=lambda kv: kv[1][0])
        dom_name, (dom_ms, dom_n) = dom
        dom_avg_ms = dom_ms / max(dom_n, 1)
        achieved = per_stage_bytes[dom_name] / (dom_avg_ms * 1e-3) / 1e9
        traffic = None
        inst = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                prof = json.load(open(tp))
                traffic = prof.get(dom_name)
                inst = prof.get("inst_executed", {}).get(dom_name)
            except Exception:
                traffic = None
        line["roofline"] = {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                            "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                            "algorithmic_bytes_per_launch": per_stage_bytes[dom_name], "avg_launch_ms": dom_avg_ms,
                            "share_of_step": dom_ms / ms_total}
        if inst is not None and clocks and clocks.get("sm_mhz") and dom_avg_ms > 0:
            # the bound that actually holds for the blend kernels: warp instructions issued per second against the issue-slot
            # peak at the SM clock measured during the run (instruction count from the committed ncu capture, time live)
            peak_issue = 148 * 4 * clocks["sm_mhz"] * 1e6
            line["roofline"]["issue"] = {"warp_inst_per_launch": inst, "achieved_per_s": inst / (dom_avg_ms * 1e-3),
                                         "peak_per_s": peak_issue, "frac": inst / (dom_avg_ms * 1e-3) / peak_issue,
                                         "note": "issue-bound kernel: HBM frac is low by construction (DESIGN.md section 5)"}
        line["roofline_step"] = {"algorithmic_bytes_per_step": step_bytes,
                                 "achieved": step_bytes / (ms_total / a.steps * 1e-3) / 1e9, "unit": "GB/s",
                                 "frac": step_bytes / (ms_total / a.steps * 1e-3) / 1e9 / peak}
        line["stage_ms_per_step"] = {k: v[0] / a.steps for k, v in stage.items()}
        line["gpu_launches"] = launches
    else:
        line["reference_kind"] = ref_kind
        line["gpu_launches"] = None
    if world == 1 and not a.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(sc, K)
        line["cpu_autograd_c1"] = cpu_autograd_c1()
    print(json.dumps(line))
    if use_dist:
        dist.destroy_process_group()
    return 0


def cpu_baseline(sc, K, max_seconds=40.0):
    """The CPU oracle port on the host cores.  Sample: the largest top-left crop of the SAME workload whose
    fwd+bwd is estimated to fit ~10-30 s (same Gaussians, same camera intrinsics, fewer pixel rows)."""
    from tests import common
    from oracle import oracle
    nthreads = oracle.max_threads()
    P, H, W = sc.P, sc.H, sc.W
    # the whole c2 frame takes ~15-30 s on 8 cores; keep full frame unless the box is small
    t0 = time.time()
    orc = common.run_oracle(sc, K, backward=True, nthreads=nthreads)
    dt = time.time() - t0
    return {"value": float(P) * H * W / dt, "unit": "Gaussian*pixel/s", "cores": nthreads, "kind": "port",
            "seconds": dt,
            "sample": f"1 fwd+bwd step of the full workload (P={P}, {H}x{W}, K={K}) with the C oracle port, {nthreads} OpenMP threads"}


def cpu_autograd_c1():
    """BASELINE.json configs[0]: SYN(10k, 256x256, K=3) forward + backward on the host through the PyTorch-autograd
    restatement (oracle/autograd_oracle.py, float64, one thread).  Reported beside the C port; not the headline."""
    try:
        from oracle import autograd_oracle
        from seganygaussians_b200 import synthetic
        P1, H1, W1, K1 = 10_000, 256, 256, 3
        sc1 = synthetic.scene(P1, H1, W1, K1)
        t0 = time.time()
        out = autograd_oracle.run_scene(sc1, K1)
        dt = time.time() - t0
        return {"value": float(P1) * H1 * W1 / dt, "unit": "Gaussian*pixel/s", "cores": 1, "kind": "port", "seconds": dt,
                "sample": f"1 fwd+bwd of SYN({P1}, {H1}x{W1}, K={K1}) (BASELINE configs[0]), torch.autograd float64, R={out.num_rendered}"}
    except Exception as e:   # pragma: no cover
        return {"value": None, "error": repr(e)}


def reference_cpu_arm(a, wl, sc):
    """Fallback of --impl reference when oracle/_ref is absent: the CPU oracle port as the reference arm."""
    cb = cpu_baseline(sc, wl["K"])
    line = {"metric": "fwd+bwd Gaussians*pixels/s @K=32", "value": cb["value"], "unit": "Gaussian*pixel/s", "n_gpus": a.gpus,
            "steps": 1, "warmup": 0, "ms_per_step": cb["seconds"] * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference", "reference_kind": "cpu-oracle-port",
            "config": {"workload": wl["desc"]}, "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "Gaussian*pixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
