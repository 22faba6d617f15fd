#!/usr/bin/env python
"""bench.py -- image-pairs/s of the dense-descriptor training hot path (fwd(A) + fwd(B) + loss + backward) at 640x480.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--pairs-per-gpu 8] [--D 3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One JSON line on stdout (rank 0).  Workload = BASELINE.json configs[1] ("batch 8 pairs, Resnet34_8s D=3,
single B200, fused fwd+loss+bwd") per GPU; N GPUs = weak scaling, 8 pairs per GPU, one NCCL gradient all-reduce
per step (configs[3] at N=8).  Inputs are synthetic (pdc_b200.synthetic), weights are the
reference's own random init.

  value        pairs/s with the step's inputs already resident in HBM (CUDA events, max over ranks)
  e2e          the same step through the reference-facing Python API starting from PINNED HOST buffers:
               H2D copies of both image batches and all index tensors and the D2H read of the loss are inside
               the timed region (what dense_correspondence/training/training.py:311-345 does per step)
  roofline     the convolution contraction kernels (forward, data-grad, weight-grad), timed individually with
               CUDA events on the launching stream during the timed region; achieved = algorithmic conv FLOPs
               (2*MACs) / summed kernel time, against the measured dense bf16 peak
  cpu_baseline the CPU oracle port of the same step on this box's host cores (bounded sample: single pairs)

``--impl reference`` times the reference's own algorithm on the host CPU (the oracle port: the reference is
Python 2 + needs its dataset stack, so it cannot run here -- see DESIGN.md) for the same metric.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

F_IMG = {3: 211.909e9, 8: 211.934e9, 16: 211.973e9}   # conv FLOPs per 640x480 image forward
CONV1_DGRAD = 1.445e9


def flops_per_pair(D, H, W):
    f = F_IMG.get(D, 211.909e9 + (D - 3) * 2 * 512 * 4800) * (H * W) / (480.0 * 640.0)
    return 2 * f + 2 * (2 * f - CONV1_DGRAD * (H * W) / (480.0 * 640.0))


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1400.0, "fallback"


class ClockSampler(object):
    """Samples SM clock / throttle reasons of one GPU every 100 ms while the timed region runs."""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self._stop = index, [], set(), threading.Event()
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        if self.nv:
            self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self.nv:
            self.t.join(timeout=2)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ---------------------------------------------------------------------------------------------------- CPU arm
_cpu_threads = [None]


def usable_cpu_threads():
    """Host threads the CPU arm can really use: affinity mask, cgroup quota, then a short calibration (a container can
    advertise 128 logical CPUs and still be throttled to a few -- 128 torch threads then run ~70x slower than 8)."""
    if _cpu_threads[0] is not None:
        return _cpu_threads[0]
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = max(1, min(n, int(int(quota) / int(period))))
    except Exception:
        pass
    from oracle.resnet34_8s_oracle import seeded_oracle
    net = seeded_oracle(D=3, seed=0).eval()
    x = torch.randn(1, 3, 240, 320)
    best, best_t = n, None
    cands = sorted({c for c in (4, 8, 16, 32, 64, n) if c <= n})
    with torch.no_grad():
        for c in cands:
            torch.set_num_threads(c)
            net(x)
            t0 = time.perf_counter(); net(x); dt = time.perf_counter() - t0
            if best_t is None or dt < best_t * 0.95:
                best, best_t = c, dt
    _cpu_threads[0] = best
    return best


def cpu_reference_step_rate(D, H, W, n_match, n_nonmatch, steps, warmup, backward=True):
    """The oracle port (plain PyTorch fp32 on the host cores): fwd(A), fwd(B), within-scene loss[, backward] on ONE pair."""
    from oracle import loss_oracle as LO
    from oracle.resnet34_8s_oracle import seeded_oracle, process_network_output
    from pdc_b200 import synthetic
    torch.set_num_threads(usable_cpu_threads())
    net = seeded_oracle(D=D, seed=0).train()
    pcl = LO.TorchPixelwiseContrastiveLoss([H, W], dict(LO.DEFAULT_LOSS_CONFIG))
    data = synthetic.make_pair_batch(1, H, W, n_match, n_nonmatch, n_nonmatch, 0, seed=1)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        net.zero_grad(set_to_none=True)
        with torch.set_grad_enabled(backward):
            ya = net(data["img_a"]); yb = net(data["img_b"])
            five = LO.batched_within_scene_loss(pcl, process_network_output(ya, 1, D, H, W),
                                                process_network_output(yb, 1, D, H, W), data)
            if backward:
                five[0].backward()
        float(five[0].detach())
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    times.sort()
    return 1.0 / times[len(times) // 2], torch.get_num_threads()


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    H, W, D = args.height, args.width, args.D
    steps, warmup = max(1, args.steps), max(1, min(args.warmup, 2))
    t0 = time.perf_counter()
    rate, cores = cpu_reference_step_rate(D, H, W, args.matches, args.non_matches, steps, warmup, backward=True)
    wall = time.perf_counter() - t0
    line = {
        "impl": "reference", "metric": "image-pairs/s (640x480, D=%d) fwd+loss+bwd" % D, "value": rate, "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": 1000.0 / rate, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: batch %d pairs/GPU, Resnet34_8s D=%d, %dx%d, %d matches + %d masked + %d background "
                               "non-matches per pair; reference arm steps over single pairs of it" %
                               (args.pairs_per_gpu, D, W, H, args.matches, args.non_matches, args.non_matches)},
        "cpu_baseline": {"value": rate, "unit": "pairs/s", "cores": cores, "kind": "port",
                         "sample": "%d timed single-pair steps (fwd A, fwd B, loss, backward) of the oracle port on the host CPU, "
                                   "median; %.1f s wall" % (steps, wall)},
        "e2e": {"value": rate, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch.distributed as dist
    import pdc_b200
    from pdc_b200 import _native as N, synthetic, loss_composer, data_parallel as DP
    from oracle import loss_oracle as LO      # only for DEFAULT_LOSS_CONFIG constants + the cpu_baseline leg

    # stdout carries exactly ONE JSON line: anything a library prints while we run (NCCL's version banner ...) goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(obj), flush=True)
        os.dup2(2, 1)

    rank, world, local_rank = DP.init_from_env()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    H, W, D, Bp = args.height, args.width, args.D, args.pairs_per_gpu
    prec_name = args.precision
    if prec_name == "auto":
        prec_name = "bf16x3" if N.lib.ddn_resnet34_8s_workspace_bytes(1, 64, 64, D, 1, N.PRECISION_BF16X3) > 0 else "fp32"
    prec = {"fp32": N.PRECISION_FP32_SIMT, "bf16x3": N.PRECISION_BF16X3, "bf16": N.PRECISION_BF16}[prec_name]

    torch.manual_seed(0)
    cfg = {"descriptor_dimension": D, "image_width": W, "image_height": H}
    dcn = pdc_b200.DenseCorrespondenceNetwork.from_config(cfg, load_stored_params=False)
    dcn.fcn.precision = prec
    DP.broadcast_parameters(dcn)
    loss_cfg = dict(LO.DEFAULT_LOSS_CONFIG)
    if args.l2_pixel_loss:
        loss_cfg["use_l2_pixel_loss_on_masked_non_matches"] = True
    pcl = pdc_b200.PixelwiseContrastiveLoss(image_shape=dcn.image_shape, config=loss_cfg)
    reducer = DP.GradientAllReducer(dcn.parameters())
    host = synthetic.make_pair_batch(Bp, H, W, args.matches, args.non_matches, args.non_matches, 0, seed=1 + rank)
    keys = [k for k, v in host.items() if v is not None]
    pinned = {k: host[k].pin_memory() for k in keys}
    resident = {k: host[k].to(dev) for k in keys}
    match_type = torch.zeros(Bp, dtype=torch.int64)          # SINGLE_OBJECT_WITHIN_SCENE, a CPU tensor like the DataLoader's
    blind = loss_composer.empty_tensor().to(dev)
    h2d_bytes = sum(pinned[k].numel() * pinned[k].element_size() for k in keys)

    def step(d):
        dcn.zero_grad(set_to_none=True)
        pa = dcn.process_network_output(dcn.forward(d["img_a"]), Bp)
        pb = dcn.process_network_output(dcn.forward(d["img_b"]), Bp)
        five = loss_composer.get_loss(pcl, match_type, pa, pb, d["matches_a"], d["matches_b"], d["masked_a"], d["masked_b"],
                                      d["background_a"], d["background_b"], blind, blind)
        five[0].backward()
        reducer()
        return five[0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(args.warmup):
        step(resident)
    barrier()

    # ---- timed region 1: inputs resident in HBM
    N.lib.ddn_profile_reset()
    N.lib.ddn_profile_enable(1)
    launches0 = N.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        barrier()
        ev0.record()
        for _ in range(args.steps):
            loss = step(resident)
        ev1.record()
        barrier()
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    launches = N.launch_count() - launches0
    N.lib.ddn_profile_enable(0)
    prof = N.profile_read()

    if args.profile_run:
        if rank == 0:
            emit({"profile_run": True, "ms_per_step_under_profiler": ms_total / args.steps, "classes": prof})
        return

    # ---- timed region 2: end to end from pinned host memory, loss read back every step.  Every step's inputs are copied
    # host->device inside the timed region (through DevicePrefetcher: the copy of step i+1 overlaps the compute of step i,
    # like a pinned-memory DataLoader would) and every step's loss is read back with .item().
    def host_batches(n):
        for _ in range(n):
            yield pinned
    for d in DP.DevicePrefetcher(host_batches(min(args.warmup, 2)), dev):
        float(step(d).item())
    barrier()
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev2.record()
    for d in DP.DevicePrefetcher(host_batches(args.steps), dev):
        last_loss = float(step(d).item())
    ev3.record()
    barrier()
    ms_e2e = max_over_ranks(ev2.elapsed_time(ev3))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pairs = world * Bp * args.steps
    value = pairs / (ms_total / 1e3)
    e2e = pairs / (ms_e2e / 1e3)
    hbm_peak, tf_peak, peak_src = measured_peaks()
    conv_ms = sum(v["ms"] for k, v in prof.items() if k.startswith("conv"))
    conv_fl = sum(v["flops"] for k, v in prof.items() if k.startswith("conv"))
    dom = max((k for k in prof if k.startswith("conv")), key=lambda k: prof[k]["ms"], default=None)
    roof = None
    if dom:
        a = prof[dom]["flops"] / (prof[dom]["ms"] * 1e-3) / 1e12 if prof[dom]["ms"] > 0 else 0.0
        mma_per_mac = {"fp32": 0, "bf16x3": 3, "bf16": 1}[prec_name]
        roof = {"bound": "tensor", "kernel": dom, "achieved": a, "peak": tf_peak, "unit": "TFLOP/s", "frac": a / tf_peak,
                # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of this kernel class (the 512->512 dilation-4 conv
                # at B=8; algorithmic 167 MB) from the committed `ncu --set full` capture profiles/r1_prof_conv_tc_final.md
                "traffic": 130.6e6 if (prec_name == "bf16x3" and Bp == 8 and H == 480 and W == 640) else None,
                "traffic_source": "profiles/r1_prof_conv_tc_final.md (largest launch of the class, bytes)",
                "issued_tensor_TFLOPs": a * mma_per_mac, "issued_frac": a * mma_per_mac / tf_peak,
                "peak_source": peak_src + " bf16_tflops_sustained (kernel timed inside a long step)",
                "launches": prof[dom]["launches"], "avg_launch_ms": prof[dom]["ms"] / max(1, prof[dom]["launches"]),
                "all_conv_achieved": (conv_fl / (conv_ms * 1e-3) / 1e12) if conv_ms > 0 else 0.0,
                "conv_share_of_step": conv_ms / (ms_total / 1.0) if ms_total > 0 else None,
                "whole_step_achieved": value * flops_per_pair(D, H, W) / world / 1e12,
                "arithmetic": {"fp32": "fp32 FFMA (CUDA cores)", "bf16x3": "bf16x3 split: 3 tensor-core MMAs per useful MAC",
                               "bf16": "single bf16 MMA"}[prec_name],
                "classes": prof}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        t0 = time.perf_counter()
        rate, cores = cpu_reference_step_rate(D, H, W, args.matches, args.non_matches, steps=3, warmup=1, backward=True)
        rate_fwd, _ = cpu_reference_step_rate(D, H, W, args.matches, args.non_matches, steps=2, warmup=0, backward=False)
        cpu = {"value": rate, "unit": "pairs/s", "cores": cores, "kind": "port",
               "sample": "oracle port, single pairs of the same workload: 3 timed fwd+loss+bwd steps (median) after 1 warm-up; "
                         "fwd+loss only = %.3f pairs/s; %.1f s of CPU wall" % (rate_fwd, time.perf_counter() - t0)}
    workload_name = ("configs[1]" if (D == 3 and Bp == 8 and args.non_matches == 1000 and not args.l2_pixel_loss) else
                     "configs[4] per-GPU shard" if (D == 8 and Bp == 4 and args.non_matches == 5000) else "custom")
    line = {
        "metric": "image-pairs/s (640x480, D=%d) fwd+loss+bwd" % D, "value": value, "unit": "pairs/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"fp32": "f32", "bf16x3": "f32 (bf16x3 split on tcgen05, fp32 accumulate)", "bf16": "bf16"}[prec_name],
        "data": "synthetic",
        "config": {"workload": workload_name + ": batch %d pairs/GPU, Resnet34_8s D=%d, %dx%d, train-mode BN, %d matches + %d masked + "
                               "%d background non-matches per pair, loss_composer.get_loss within-scene, no optimizer step" %
                               (Bp, D, W, H, args.matches, args.non_matches, args.non_matches),
                   "l2_pixel_loss_on_masked_non_matches": bool(args.l2_pixel_loss), "global_batch_pairs": world * Bp, "parallelism": "dp%d" % world, "precision": prec_name,
                   "l2": "inputs+activations touched per step (~%.1f GB) are far larger than the 126 MB L2; no explicit flush" %
                         (N.lib.ddn_resnet34_8s_workspace_bytes(Bp, H, W, D, 1, prec) * 2 / 1e9)},
        "clocks": clk.summary(),
        "e2e": {"value": e2e, "unit": "pairs/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps, "last_loss": last_loss},
        "gpu_launches": launches,
        "roofline": roof,
        "cpu_baseline": cpu,
        "loss": float(loss.item()),
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs-per-gpu", type=int, default=8)
    ap.add_argument("--D", type=int, default=3)
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--matches", type=int, default=1000)
    ap.add_argument("--non-matches", type=int, default=1000)
    ap.add_argument("--precision", default="auto", choices=["auto", "fp32", "bf16x3", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--l2-pixel-loss", action="store_true",
                    help="configs[4] variant: use_l2_pixel_loss_on_masked_non_matches=True (M_pixel=50)")
    ap.add_argument("--profile-run", action="store_true",
                    help="short run for ncu: 1 warm-up + --steps timed steps, no e2e / cpu legs (numbers printed are NOT bench values)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            sys.stderr.write("bench.py: --gpus %d needs a torchrun launch (WORLD_SIZE=%d); see the module docstring\n" % (args.gpus, world))
            sys.exit(2)
    if args.profile_run:
        args.warmup, args.no_cpu_baseline = 1, True
    elif args.warmup < 3:
        args.warmup = 3
    run_ours(args)


if __name__ == "__main__":
    main()
