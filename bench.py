#!/usr/bin/env python
"""Headline benchmark: SNN-detector training throughput (images/s) on B200 through the libsnnb200 hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W]                 # our arm (1 GPU, or under torchrun: N ranks)
    python bench.py --impl reference [--steps K] [--warmup W]           # CPU arm: oracle port of the reference path
    python bench.py --microbench lif                                    # BASELINE.json configs[3] LIF sweep (GB/s)

A "step" is one training step of the reference (train.py:58-80): T-frame unroll with state carry, detection loss on
the last frame, backward, clip_grad_norm_(10), AdamW, OneCycleLR -- on one synthetic batch.  Workload at every N is
BASELINE.json configs[1] per GPU (default SNN detector, T=4, batch 64, 256x256 frames; weak scaling);
``--config 3`` selects configs[2] (T=8, 512x512).

Prints ONE JSON line (rank 0).  `value` = device-resident throughput (inputs already in HBM), `e2e` = the same metric
through the public API with pinned HOST buffers (H2D of the frames + labels and a D2H read of the loss every step),
`roofline` = the dominant kernel timed live with CUDA events inside the timed region, `cpu_baseline` = the oracle
port of the reference path timed on this box's host cores (N=1 only).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

HYP = {"box": 7.5, "cls": 1.0, "dfl": 2.5, "reg_max": 16}      # reference config.yaml:33-37
NUM_CLASSES = 8                                                 # config.yaml:29
MAX_LR, WEIGHT_DECAY = 1e-4, 5e-4                               # config.yaml:23-24


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 3], help="BASELINE.json configs index + 1")
    ap.add_argument("--batch", type=int, default=None, help="per-GPU batch (default: the config's)")
    ap.add_argument("--neuron", default="lif", choices=["lif", "silu"])
    ap.add_argument("--microbench", default=None, choices=["lif"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-profile", action="store_true", help="skip the per-kernel CUDA-event accounting")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying one CUDA graph")
    ap.add_argument("--ref-batch", type=int, default=2, help="sequences per step of the CPU arm (bounded sample)")
    return ap.parse_args()


def workload(args):
    if args.config == 1:
        B, T, HW = 2, 4, 256
    elif args.config == 2:
        B, T, HW = 64, 4, 256
    else:
        B, T, HW = 16, 8, 512
    if args.batch:
        B = args.batch
    return B, T, HW


def workload_name(args, B, T, HW):
    return (f"BASELINE.json configs[{args.config - 1}]: default SNN detector ({args.neuron} neurons), T={T}, batch {B}/GPU, "
            f"synthetic {HW}x{HW} RGB frames, one train step (fwd + loss + bwd + clip + AdamW + OneCycle)")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback (B200_PROFILING.md)")


# ------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, pw = [], None, set(), []
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1]); pw.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [s for s, p in zip(sm, pw) if p > 250.0] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "power_w_max": max(pw) if pw else None, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm: oracle port of the reference path (test infrastructure used as the measured baseline)
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(T, HW, batch, steps, warmup, neuron):
    import torch
    from oracle import model_oracle as MO
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(42)
    model = MO.OracleYOLOTemporalUNet(num_classes=NUM_CLASSES, hyp=HYP, neuron=neuron)
    MO.initialize_model_oracle(model)
    model.train()
    loss_fn, opt, sched = MO.make_reference_trainer(model, total_steps=max(1000, steps + warmup + 1), max_lr=MAX_LR,
                                                    weight_decay=WEIGHT_DECAY)
    frames, labels = MO.synthetic_batch(batch, T, HW, HW, nc=NUM_CLASSES, seed=42)
    for _ in range(warmup):
        MO.reference_train_step(model, loss_fn, opt, sched, frames, labels)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        MO.reference_train_step(model, loss_fn, opt, sched, frames, labels)
        times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return dict(value=batch / sec, sec_per_step=sec, cores=cores, threads=torch.get_num_threads(),
                sample=f"{steps} timed + {warmup} warm-up training steps of {batch} sequences (T={T}, {HW}x{HW}) = a "
                       f"{batch}-sequence slice of the workload's batch; oracle port (PyTorch fp32 eager, {cores} threads)")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B, T, HW = workload(args)
    r = cpu_reference_run(T, HW, args.ref_batch, args.steps, args.warmup, args.neuron)
    line = {
        "impl": "reference", "metric": "train images/sec", "value": r["value"], "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["sec_per_step"] * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args, B, T, HW), "per_gpu_batch": B, "T": T, "cpu_sample_batch": args.ref_batch},
        "cpu_baseline": {"value": r["value"], "unit": "images/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def summarize_profile(records, steps, pk):
    """records: (name, work, ev0, ev1) of every ABI call in the timed region -> per-entry-point totals."""
    agg = {}
    for name, work, e0, e1 in records:
        ms = e0.elapsed_time(e1)
        a = agg.setdefault(name, dict(ms=0.0, calls=0, flop=0.0, byte=0.0))
        a["ms"] += ms
        a["calls"] += 1
        if work:
            a[work[0]] += work[1]
    shapes = {}
    for name, work, e0, e1 in records:
        if work and len(work) > 2:
            a = shapes.setdefault((name, work[2]), dict(ms=0.0, calls=0, amount=0.0, kind=work[0]))
            a["ms"] += e0.elapsed_time(e1); a["calls"] += 1; a["amount"] += work[1]
    by_shape = []
    for (name, tag), a in sorted(shapes.items(), key=lambda kv: -kv[1]["ms"])[:40]:
        rate = a["amount"] / (a["ms"] * 1e-3)
        by_shape.append({"kernel": name, "shape": tag, "ms_per_step": round(a["ms"] / steps, 4), "calls_per_step": a["calls"] / steps,
                         ("tflops" if a["kind"] == "flop" else "gbs"): round(rate / (1e12 if a["kind"] == "flop" else 1e9), 1)})
    summarize_profile.by_shape = by_shape
    out = {}
    for name, a in agg.items():
        d = dict(ms_per_step=a["ms"] / steps, calls_per_step=a["calls"] / steps)
        if a["flop"]:
            d["tflops"] = a["flop"] / (a["ms"] * 1e-3) / 1e12
            d["frac_of_bf16_sustained"] = d["tflops"] / pk["tf_sust"]
        if a["byte"]:
            d["gbs"] = a["byte"] / (a["ms"] * 1e-3) / 1e9
            d["frac_of_hbm"] = d["gbs"] / pk["hbm"]
        out[name] = d
    return out, agg


def run_ours(args):
    import torch
    import torch.distributed as dist
    from snn_object_detectionddp_b200 import _lib
    from snn_object_detectionddp_b200.model import YOLOTemporalUNet
    from snn_object_detectionddp_b200.trainer import Trainer
    from snn_object_detectionddp_b200.weight_initialization import initialize_model
    from snn_object_detectionddp_b200.data import synthetic_batch

    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.lib()        # fail loudly if libsnnb200.so is missing: there is no fallback
    B, T, HW = workload(args)
    pk = peaks()

    torch.manual_seed(42)
    model = YOLOTemporalUNet(num_classes=NUM_CLASSES, yolo_model_name="yolo11m.pt", use_conv_lstm=True, hyp=HYP,
                             neuron=args.neuron)
    initialize_model(model)
    trainer = Trainer(model, max_lr=MAX_LR, weight_decay=WEIGHT_DECAY, total_steps=1000, device=dev)
    frames_cpu, labels_cpu = synthetic_batch(B, T, HW, HW, nc=NUM_CLASSES, seed=42 + rank)
    frames = frames_cpu.to(dev)
    MAXB = 8                                              # synthetic labels: 0-7 boxes per sample
    batch_dev = {"padded": tuple(t.to(dev) for t in trainer.prepare_batch(labels_cpu, B, max_boxes=MAXB)["padded"])}
    graphed = not args.no_graph
    step_fn = trainer.train_step_graphed if graphed else trainer.train_step

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput ----------------
    launches0 = _lib.launch_count
    for _ in range(max(args.warmup, 3)):                  # >= 3: two eager steps, then the graph capture + first replay
        step_fn(frames, batch_dev)
    barrier()
    calls_per_step = None
    if graphed:                                           # ABI calls recorded into the graph = launches per replay
        l0 = _lib.launch_count
        trainer.train_step(frames, batch_dev)
        calls_per_step = _lib.launch_count - l0
        barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    records = None if (args.no_profile or graphed) else []
    _lib.profile = records
    launches0 = _lib.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        loss, items = step_fn(frames, batch_dev)
    ev1.record()
    barrier()
    _lib.profile = None
    launches = calls_per_step * args.steps if graphed else _lib.launch_count - launches0
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms)
    last_loss = [float(v) for v in items]

    # ---------------- end to end through the public API with host buffers ----------------
    e2e = None
    if not args.no_e2e:
        # host frames as the dataset decodes them: uint8 RGB (reference dataset.py:139-152 converts to fp32 / 255 on the HOST and
        # ships 4x the bytes); the division runs in the frame packer kernel, bit-identical to the host's
        frames_u8 = (frames_cpu * 255.0).round().clamp_(0, 255).to(torch.uint8)
        from snn_object_detectionddp_b200.data import DevicePrefetcher
        host_batches = [(frames_u8.clone().pin_memory(),
                         tuple(t.pin_memory() for t in trainer.prepare_batch(labels_cpu, B, max_boxes=MAXB)["padded"])) for _ in range(2)]
        loss_host = torch.zeros(args.steps + args.warmup, 3).pin_memory()
        pf = DevicePrefetcher(dev)

        def e2e_steps(n, base):
            pf.stage(*host_batches[0])
            for i in range(n):
                frames_d, padded_d = pf.take()
                if i + 1 < n:
                    pf.stage(*host_batches[(i + 1) % 2])        # prefetch the next batch while this step computes
                _, it = step_fn(frames_d, {"padded": padded_d})
                pf.release()
                loss_host[base + i].copy_(it, non_blocking=True)     # D2H read of the step's result

        e2e_steps(args.warmup, 0)
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        e2e_steps(args.steps, args.warmup)
        t1.record()
        barrier()
        ems = torch.tensor([t0.elapsed_time(t1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ems, op=dist.ReduceOp.MAX)
        h2d_bytes = pf.bytes_per_batch()
        e2e = {"value": world * B * args.steps / (float(ems) * 1e-3), "unit": "images/s", "h2d_bytes_per_step": h2d_bytes,
               "d2h_bytes_per_step": 12, "ms_per_step": float(ems) / args.steps,
               "api": ("Trainer.train_step_graphed" if graphed else "Trainer.train_step")
                      + " fed by data.DevicePrefetcher: double-buffered pinned host frames (uint8 [B,T,3,H,W], /255 on the device)"
                        " + padded labels, copied on a side stream"}

    clocks = sampler.stop() if rank == 0 else None          # sampled over both timed regions (device-resident + e2e)

    # ---------------- per-kernel accounting + roofline of the dominant kernel ----------------
    roofline, kernels_summary = None, None
    if graphed and not args.no_profile:
        # a replayed graph has no per-launch events: time the SAME kernels in an instrumented eager pass of the same
        # K steps right after the timed region (kernel durations do not depend on how they were launched)
        records = []
        _lib.profile = records
        pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        pe0.record()
        for _ in range(args.steps):
            trainer.train_step(frames, batch_dev)
        pe1.record()
        barrier()
        _lib.profile = None
        eager_ms_total = pe0.elapsed_time(pe1)
    else:
        eager_ms_total = ms_total
    if records:
        kernels_summary, agg = summarize_profile(records, args.steps, pk)
        top = max(agg.items(), key=lambda kv: kv[1]["ms"])
        name, a = top
        if a["flop"]:
            ach = a["flop"] / (a["ms"] * 1e-3) / 1e12
            roofline = {"kernel": name, "bound": "tensor", "achieved": ach, "peak": pk["tf_sust"], "unit": "TFLOP/s",
                        "frac": ach / pk["tf_sust"], "frac_of_nominal_2250": ach / 2250.0, "traffic": None,
                        "peak_source": pk["src"] + ", sustained bf16",
                        "share_of_step": a["ms"] / ms_total, "launches_per_step": a["calls"] / args.steps}
        else:
            ach = a["byte"] / (a["ms"] * 1e-3) / 1e9
            roofline = {"kernel": name, "bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s",
                        "frac": ach / pk["hbm"], "frac_of_nominal_8000": ach / 8000.0, "traffic": None, "peak_source": pk["src"],
                        "share_of_step": a["ms"] / ms_total, "launches_per_step": a["calls"] / args.steps}
        tr = os.path.join(ROOT, "profiles", "traffic.json")       # dram bytes per launch from the committed ncu capture
        if os.path.isfile(tr):
            roofline["traffic"] = json.load(open(tr)).get(name)

    def shutdown():
        """Leave without waiting on NCCL teardown: the captured graph still references the communicator's kernels and
        destroy_process_group() was seen to block for minutes on the 2-GPU box after the result line was out."""
        sys.stdout.flush()
        sys.stderr.flush()
        if world > 1:
            trainer._graph = None
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()
            os._exit(0)

    if world > 1:
        dist.barrier()
    if rank != 0:
        shutdown()
        return

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(T, HW, args.ref_batch, 3, 1, args.neuron)
        cpu_baseline = {"value": r["value"], "unit": "images/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]}

    line = {
        "metric": "train images/sec", "value": world * B * args.steps / (ms_total * 1e-3), "unit": "images/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": workload_name(args, B, T, HW),
                   "per_gpu_batch": B, "T": T, "frames_per_s": world * B * T * args.steps / (ms_total * 1e-3),
                   "l2": "per-step working set (240 MB bf16 weights + GBs of activations) exceeds the 126 MB L2; no explicit flush",
                   "parallelism": f"dp{world}", "feature_extractor": "stand-in frozen pyramid (YOLO11m weights unobtainable offline)"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu_baseline,
        "cuda_graph_active": bool(graphed and trainer._graph is not None and not trainer._graph_failed),
        "kernels": kernels_summary, "kernels_by_shape": getattr(summarize_profile, "by_shape", None), "last_loss_items": last_loss, "cuda_graph": graphed,
        "kernel_ms_sum_per_step": (sum(v["ms_per_step"] for v in kernels_summary.values()) if kernels_summary else None),
        "eager_instrumented_ms_per_step": (eager_ms_total / args.steps if kernels_summary else None),
    }
    print(json.dumps(line), flush=True)
    shutdown()


# ------------------------------------------------------------------------------------------------
# LIF microbench (BASELINE.json configs[3])
# ------------------------------------------------------------------------------------------------
def run_lif_microbench(args):
    import torch
    from snn_object_detectionddp_b200 import kernels as K
    pk = peaks()
    torch.cuda.set_device(0)
    rows = []
    for T in (4, 8, 16):
        for C, HWs in ((64, 64), (128, 64), (128, 32), (256, 32), (256, 16), (512, 16)):
            per_img = C * HWs * HWs
            Bn = max(1, (1 << 30) // (4 * T * per_img))          # fp32 input tensor >= 1 GiB (L2 is 126 MB)
            y = torch.randn(T * Bn, HWs, HWs, C, device="cuda")
            scale, shift = torch.ones(T, C, device="cuda"), torch.full((T, C), 0.2, device="cuda")
            mean, invstd = torch.zeros(T, C, device="cuda"), torch.ones(T, C, device="cuda")
            gs = torch.randn(T * Bn, HWs, HWs, C, device="cuda").to(torch.bfloat16)
            n = y.numel()

            def timeit(fn, reps=5):
                fn(); torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    fn()
                e1.record(); torch.cuda.synchronize()
                return e0.elapsed_time(e1) / reps

            bnb = torch.full((C,), 0.2, device="cuda")
            dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
            from snn_object_detectionddp_b200._lib import call, ptr, stream_ptr
            red = torch.zeros(T, 2, C, device="cuda")
            dy = torch.empty_like(gs)
            P = n // (T * C)
            a = (ptr(y), ptr(scale), ptr(shift), ptr(mean), ptr(invstd), ptr(bnb), None, ptr(gs), None, ptr(red), ptr(dy), None,
                 ptr(dg), ptr(db), T, P, C, 0.5, 1.0, 2.0)
            f_ms = timeit(lambda: K.bn_act_fwd(0, y, scale, shift, T))
            r_ms = timeit(lambda: call("snn_bn_act_bwd2", 0, 0, *a, stream_ptr()))
            d_ms = timeit(lambda: call("snn_bn_act_bwd2", 1, 0, *a, stream_ptr()))
            be_ms = timeit(lambda: K.bn_act_bwd(0, False, y, scale, shift, mean, invstd, gs, T))
            # algorithmic bytes / neuron-timestep: fwd 4+2+1/8; train bwd pass 1 (reductions) 4+2, pass 2 (dx) 4+2+2;
            # frozen-statistics bwd 4+2+2
            rows.append(dict(T=T, C=C, HW=HWs, B=Bn, fwd_gbs=n * 6.125 / f_ms / 1e6, bwd_reduce_gbs=n * 6.0 / r_ms / 1e6,
                             bwd_dx_gbs=n * 8.0 / d_ms / 1e6, bwd_train_gbs=n * 14.0 / (r_ms + d_ms) / 1e6,
                             bwd_eval_gbs=n * 8.0 / be_ms / 1e6, fwd_ms=f_ms, bwd_reduce_ms=r_ms, bwd_dx_ms=d_ms))
            del y, gs, dy
    for r in rows:
        r["fwd_frac"], r["bwd_frac"] = r["fwd_gbs"] / pk["hbm"], r["bwd_train_gbs"] / pk["hbm"]
        r["fwd_frac_of_nominal_8000"], r["bwd_frac_of_nominal_8000"] = r["fwd_gbs"] / 8000.0, r["bwd_train_gbs"] / 8000.0
    print(json.dumps({"microbench": "lif", "hbm_peak_gbs": pk["hbm"], "peak_source": pk["src"], "rows": rows}), flush=True)


if __name__ == "__main__":
    a = parse()
    if a.microbench == "lif":
        run_lif_microbench(a)
    elif a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
