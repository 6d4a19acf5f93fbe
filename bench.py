#!/usr/bin/env python
"""bench.py - 256^2 perceptual-loss training images/sec (BASELINE.json metric) on N B200s of one node.

  python bench.py --gpus 1 --steps K --warmup W              # our arm (libast_b200.so kernels)
  torchrun --nproc-per-node N ... bench.py --gpus N ...       # data-parallel, one rank per GPU, NCCL allreduce
  python bench.py --impl reference ...                        # the reference's CPU step (oracle port) on host cores

A "step" is one full optimisation step of train_cnn.py:295-334 (TransformerNet fwd/bwd, VGG taps, Grams,
style/content MSE, gradient all-reduce, Adam) on one synthetic batch.  Workload at every N: BASELINE
configs[1] - batch 32 per GPU at 256x256, bf16-in/fp32-accumulate ("fast" precision) - weak scaling.
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic work per training image at 256^2 (scales with H*W)
GF_PER_IMG = {"total": 137.86, "conv_gather": 16.803 + 15.784 + 36.465 + 36.465 + 12.306 + 2.147,
              "wgrad_gather": 16.803, "gram": 1.082}
IN_ELEMS_PER_IMG = 13.107e6   # elements through the 17 InstanceNorm layers at 256^2


def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "src": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.proc, self.idx = None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self, t_begin=None, t_end=None):
        """Median SM clock / throttle reasons of the samples whose timestamp falls inside [t_begin, t_end] (the sampler
        is started before the warm-up so that it is already streaming when the timed region begins)."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        import datetime
        sm, mx, reasons, power = [], None, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                clk, mxv = float(f[1]), float(f[2])
            except ValueError:
                continue
            if t_begin is not None and not (t_begin - 0.03 <= ts <= t_end + 0.03):
                continue
            sm.append(clk); mx = mxv
            try:
                power.append(float(f[3]))
            except ValueError:
                pass
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_min_mhz": sm[0] if sm else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_max": max(power) if power else None}


def _finish(world):
    """Multi-rank runs leave through os._exit: tearing down an NCCL communicator whose collectives were captured in a
    (still alive) CUDA graph blocked both ranks for minutes at interpreter exit on the B200 boxes."""
    if world > 1:
        sys.stdout.flush(); sys.stderr.flush()
        os._exit(0)


def cpu_reference_step_rate(batch, size, steps, warmup, threads):
    """The reference's training step (oracle port of train_cnn.py:295-334 + Adam) on the host cores."""
    import torch
    from oracle import port, weights
    torch.set_num_threads(threads)
    tsd = port.make_leaf(weights.transfer_state_dict(2), torch.float32)
    vsd = weights.vgg_state_dict(2)
    style = port.style_grams_single(weights.style_image(size, 2), vsd, batch)
    state, params = {}, tsd
    times = []
    for it in range(warmup + steps):
        content = weights.content_batch(batch, size, 2, step=it)
        t0 = time.perf_counter()
        leaf = port.make_leaf(params, torch.float32)
        out = port.training_step(leaf, vsd, content, style)
        params = port.adam_l2_step({k: v.detach() for k, v in leaf.items()}, out["grads"], state, 0.0024, it + 1)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return batch * len(times) / sum(times), sum(times) / len(times) * 1e3


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    batch = 4
    rate, ms = cpu_reference_step_rate(batch, args.size, args.steps, args.warmup, threads)
    sample = f"B={batch} images of {args.size}x{args.size} per step, fp32, torch CPU ops, {threads} threads"
    line = {"impl": "reference", "metric": "train_images_per_sec_256", "value": rate, "unit": "images/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"perceptual-loss training step, {args.size}x{args.size}, CPU sample of B={batch} "
                                   "(BASELINE configs[1] is B=32/GPU)"},
            "cpu_baseline": {"value": rate, "unit": "images/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_ours(args):
    # keep stdout clean for the ONE JSON line: libraries (e.g. "NCCL version ...") write to fd 1 during init
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    import artist_style_transfer_b200 as ast
    from artist_style_transfer_b200 import _lib, ops

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback for the product path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    B, S, K, W = args.batch, args.size, args.steps, args.warmup

    torch.manual_seed(2)                                   # identical replicas on every rank (SEED, train_cnn.py:44)
    net = ast.StyleTransfer(device=dev, precision=args.precision)
    vgg = ast.VGG16(vgg_path=None, precision=args.precision).to(dev)
    gen = torch.Generator(device=dev).manual_seed(2 + 7919 * rank)
    style_img = torch.randint(0, 256, (3, S, S), device=dev, generator=torch.Generator(device=dev).manual_seed(2)).float()
    style = ast.style_grams_single(vgg, style_img, B)
    trainer = ast.PerceptualTrainer(net, vgg, style, cuda_graph=not args.no_graph)
    nbuf = 4
    batches = [torch.randint(0, 256, (B, 3, S, S), device=dev, generator=gen).float() for _ in range(nbuf)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for i in range(W + (4 if not args.no_graph else 0)):   # graph mode: 3 eager steps + capture happen before timing
        trainer.step(batches[i % nbuf])
    barrier()
    t_begin = time.time()
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        losses = trainer.step(batches[i % nbuf])
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop(t_begin, time.time()) if rank == 0 else None
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = B * world * K / (ms / 1e3)

    # ---- e2e: host buffers -> H2D copy -> step -> D2H loss read, every step (same workload)
    host = [torch.randint(0, 256, (B, 3, S, S)).float().pin_memory() for _ in range(2)]
    loss_host = torch.zeros(3, dtype=torch.float32).pin_memory()
    ke = max(2, min(K, 5))
    barrier()
    e0.record()
    nxt = trainer.prefetch(host[0])                         # H2D of step 0's inputs (inside the timed region)
    for i in range(ke):
        x = nxt
        if i + 1 < ke:                                      # public API: the next batch's H2D copy runs on a side stream
            nxt = trainer.prefetch(host[(i + 1) % 2])       # while this step computes; every step still copies its inputs
        c, s, tot = trainer.step(x)
        loss_host.copy_(torch.stack([c, s, tot]), non_blocking=True)
        torch.cuda.current_stream().synchronize()           # the caller reads the loss
    e1.record()
    barrier()
    te = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = B * world * ke / (float(te.item()) / 1e3)

    # ---- per-kernel-family device time (CUDA events on the launching stream), 2 instrumented steps
    lc0 = _lib.launch_count()
    ops.profile_begin()
    kp = 2
    for i in range(kp):
        trainer._eager_step(batches[i % nbuf])          # per-op events need eager launches (not a graph replay)
    prof = ops.profile_end()
    if not args.no_graph:        # replays do not pass through the library's launch counter: use the eager count
        launches = (_lib.launch_count() - lc0) // kp * K
    fam = {k: {"ms_per_step": v[0] / kp, "launches_per_step": v[1] / kp} for k, v in prof.items()}
    scale = (S / 256.0) ** 2
    peaks = read_peaks()
    conv_ms = sum(fam[k]["ms_per_step"] for k in ("conv_gather_tc", "conv_gather_simt") if k in fam)
    conv_launches = sum(fam[k]["launches_per_step"] for k in ("conv_gather_tc", "conv_gather_simt") if k in fam)
    conv_tf = GF_PER_IMG["conv_gather"] * scale * B / conv_ms            # GF/ms == TF/s
    peak_tf = peaks["bf16_tflops_sustained"]
    # TF32 peak is not in MEASURED_PEAKS.json: measure it the same way (cuBLAS 8192^3, best of 5).
    # 42 % of the conv FLOPs of this mode run in TF32 (VGG forward + Gram backward), the rest in bf16.
    tf32_peak = None
    if rank == 0 and args.precision == "fast":
        old = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        a_ = torch.randn(8192, 8192, device=dev); b_ = torch.randn(8192, 8192, device=dev)
        best = 1e9
        for _ in range(6):
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record(); torch.matmul(a_, b_); t1.record(); torch.cuda.synchronize()
            best = min(best, t0.elapsed_time(t1))
        tf32_peak = 2 * 8192 ** 3 / best / 1e9
        torch.backends.cuda.matmul.allow_tf32 = old
        del a_, b_
    # TF32: VGG forward on the generated batch (36.465) and on the content batch (12.306) + the Gram backward (2.147);
    # bf16: TransformerNet forward / dgrad and, since the bf16 gradient chain, the VGG dgrad (36.465)
    tf32_share = (36.465 + 12.306 + 2.147) / GF_PER_IMG["conv_gather"]
    mix_peak = None if not tf32_peak else 1.0 / (tf32_share / tf32_peak + (1 - tf32_share) / peaks["bf16_tflops"])
    roofline = {"kernel": "conv_gather (all conv fwd/dgrad launches of one step: conv_px_kernel + conv_tc_kernel + conv_ws_kernel)",
                "bound": "tensor", "achieved": conv_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": conv_tf / peak_tf,
                "traffic": 491.8e6, "traffic_note": "dram read+write bytes of one conv_px<tf32> launch (128->128 @128^2, "
                "B=32) from profiles/r01f_conv_px_ncu_full_excerpt.csv; algorithmic bytes of that launch: 537 MB",
                "peak_source": f"{peaks['src']} bf16 sustained", "launch_ms_avg": conv_ms / max(1.0, conv_launches),
                "tf32_tflops_measured": tf32_peak, "tf32_flop_share": tf32_share,
                "frac_of_precision_mix_peak": None if not mix_peak else conv_tf / mix_peak}
    in_ms = fam.get("instnorm", {"ms_per_step": float("nan")})["ms_per_step"]
    esz = 2 if args.precision == "fast" else 4
    in_gb = IN_ELEMS_PER_IMG * scale * B * esz * 5 / 1e9                 # fwd 1R+1W, bwd 2R+1W
    # the kernels physically move 7 passes (apply 1R+1W; backward statistics 2R, backward apply 2R+1W; the forward
    # statistics ride in the conv epilogue) where 5 are algorithmic
    roofline_in = {"kernel": "instnorm (in_apply_staged fwd; in_bwd_stats_staged + in_bwd_apply_staged bwd)", "bound": "hbm",
                   "achieved": in_gb / (in_ms / 1e3), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                   "frac": in_gb / (in_ms / 1e3) / peaks["hbm_gbs"], "traffic": None,
                   "achieved_physical": in_gb * 7 / 5 / (in_ms / 1e3),
                   "frac_physical": in_gb * 7 / 5 / (in_ms / 1e3) / peaks["hbm_gbs"]}

    if world > 1:                      # everyone is done with the GPU work; from here on only rank 0 has something to do
        torch.cuda.synchronize()
        dist.barrier()
    if rank != 0:
        _finish(world)
        return
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        rate, _ = cpu_reference_step_rate(4, S, 3, 1, threads)
        cpu_baseline = {"value": rate, "unit": "images/s", "cores": threads, "kind": "port",
                        "sample": f"1 warm-up + 3 timed steps of B=4 at {S}x{S}, fp32 oracle port of train_cnn.py:295-334 + Adam"}
    line = {
        "metric": "train_images_per_sec_256", "value": value, "unit": "images/s", "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.precision == "fast" else "f32", "data": "synthetic",
        "config": {"workload": f"BASELINE configs[1]: perceptual-loss training step, {S}x{S}, batch {B}/GPU, "
                               f"precision={args.precision}, random-init TransformerNet+VGG16, Adam",
                   "global_batch": B * world, "parallelism": f"dp{world}", "cuda_graph": not args.no_graph,
                   "l2": "per-step working set (GBs of activations) >> 126 MB L2; inputs rotate over 4 buffers"},
        "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": B * 3 * S * S * 4,
                "d2h_bytes_per_step": 12, "steps": ke},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "roofline_instnorm": roofline_in,
        "kernel_families": fam, "cpu_baseline": cpu_baseline,
        "losses_last_step": [float(x) for x in losses],
        "gflop_per_image": GF_PER_IMG["total"] * scale,
        "tensor_frac_whole_step": GF_PER_IMG["total"] * scale * B * K / ms / peak_tf,
    }
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    _finish(world)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="fast", choices=["fast", "fp32"])
    ap.add_argument("--batch", type=int, default=32, help="images per GPU per step")
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying a CUDA graph")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
