#!/usr/bin/env python
"""bench.py -- images/sec of the Darknet hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

N=1 workload = BASELINE.json configs[1]: yolov3.cfg (Darknet-53) inference, batch 32, 640x640, fp16 fused conv
path, synthetic images, random-init weights.  A "step" is one forward (stem conv -> 74 tcgen05 convs -> YOLO
decode) of one batch.  N>1: independent replicas, one batch per GPU per step, no collective on the data path
("scaling": "weak").

One JSON line on stdout (rank 0):
  value  : images/s with the fp32 batch already resident in HBM (CUDA events, max over ranks)
  e2e    : images/s through models.Darknet.__call__ with pinned HOST uint8 images: H2D copy, /256, forward,
           D2H of the top-objectness detection row per image -- all inside the timed region
  roofline: aggregate over the tcgen05 conv launches of one step (the dominant kernel family), timed live
  cpu_baseline: the oracle (CPU restatement of the reference's PyTorch path) on this box's host cores, bounded sample
--impl reference: the reference arm = that same CPU path with all host threads (the reference itself is Python and
cannot travel to the GPU box; oracle/ is pinned against it by tests/test_oracle_golden.py).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "yolov3v4-modelcompression-multidatasettraining-multibackbone_b200")
sys.path.insert(0, PKG)

import torch  # noqa: E402

MODEL = "yolov3"
BATCH = 32
SIZE = 640
# conv MACs / image, yolov3 @640 (recomputed from the plan at run time as a cross-check)
FLOPS_PER_IMAGE = 155.89e9


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "tflops": d["bf16_tflops_sustained"], "tflops_burst": d["bf16_tflops"],
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "tflops": 1400.0, "tflops_burst": 1590.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        for ln in self.lines:
            f = [c.strip() for c in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def build_model(device):
    import models
    from b200yolo import cfggen
    cfg_dir = os.path.join(ROOT, "gpurun_out", "cfg") if os.access(ROOT, os.W_OK) else "/tmp/b2y_cfg"
    path = cfggen.write_cfg(MODEL, cfg_dir)
    torch.manual_seed(0)
    m = models.Darknet(path, img_size=(SIZE, SIZE))
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():  # non-degenerate BN statistics for eval
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.running_var.copy_(torch.rand(mod.running_var.shape, generator=g) * 0.4 + 0.8)
                mod.running_mean.copy_(torch.randn(mod.running_mean.shape, generator=g) * 0.1)
    m = m.to(device).eval()
    m.static_outputs = True   # return the engine's output buffers (no per-step clone)
    return m


def cpu_reference_rate(max_seconds=20.0, threads=None, batch=2, max_iters=10):
    """The reference's CPU PyTorch path, restated in oracle/ (pinned to the reference by tests/golden):
    yolov3 eval forward at 640x640 on the host cores, bounded sample."""
    sys.path.insert(0, ROOT)
    from oracle import darknet_oracle as orc
    import models
    from b200yolo import cfggen
    from utils.parse_config import parse_model_cfg_text
    try:
        avail = len(os.sched_getaffinity(0))      # cores this process may be scheduled on
    except AttributeError:
        avail = os.cpu_count()
    try:                                           # ... and the cgroup CPU quota, if any (cpu.max = "<quota> <period>")
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if q != "max":
            avail = max(1, min(avail, int(float(q) / float(per) + 0.5)))
    except (OSError, ValueError):
        pass
    defs = parse_model_cfg_text(cfggen.cfg_text(MODEL))[1:]
    path = cfggen.write_cfg(MODEL, "/tmp/b2y_cfg_cpu")
    sd = orc.synth_state_dict(models.Darknet(path).state_dict(), 0)
    if threads is None:
        # "all the host threads it can use": more threads than physically free cores make torch's CPU convs slower,
        # so take the fastest of a few thread counts on a small probe forward
        best = None
        probe = orc.synth_images(1, 256, 256, seed=1)
        for t in sorted({avail, min(avail, 64), min(avail, 32), min(avail, 16), min(avail, 8)}, reverse=True):
            torch.set_num_threads(t)
            with torch.no_grad():
                orc.darknet_forward(defs, sd, probe, MODEL)
                t0 = time.time()
                orc.darknet_forward(defs, sd, probe, MODEL)
                dt = time.time() - t0
            if best is None or dt < best[0]:
                best = (dt, t)
        threads = best[1]
    torch.set_num_threads(threads)
    bs = batch
    x = orc.synth_images(bs, SIZE, SIZE, seed=0)
    with torch.no_grad():
        t0 = time.time()
        orc.darknet_forward(defs, sd, x, MODEL)          # warm-up (also sizes the sample)
        warm = time.time() - t0
        iters = max(1, min(max_iters, int(max_seconds / max(warm, 1e-3))))
        t0 = time.time()
        for _ in range(iters):
            orc.darknet_forward(defs, sd, x, MODEL)
        dt = (time.time() - t0) / iters
    return {"value": bs / dt, "unit": "images/s", "cores": threads, "kind": "port",
            "sample": "%d x yolov3 eval forward, batch %d, 640x640, fp32 torch CPU (oracle/darknet_oracle.py)"
                      % (iters, bs), "ms_per_step": dt * 1e3}, bs


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    # a "step" of the reference arm = one yolov3 640x640 eval forward of ONE image on all host cores; the number of
    # timed steps is capped so that the whole run stays within ~2 minutes (stated in cpu_baseline.sample)
    base, bs = cpu_reference_rate(max_seconds=100.0, batch=1, max_iters=max(1, args.steps))
    line = {"impl": "reference", "metric": "images/sec (640x640) yolov3 inference", "value": base["value"],
            "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": base["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "yolov3.cfg Darknet-53 inference 640x640 (BASELINE configs[1]); CPU sample batch %d"
                                   % bs},
            "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": base["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def ncu_conv_traffic():
    """DRAM bytes (read + write) of the tcgen05 conv launches of ONE forward, from the committed ncu launch list
    (profiles/r01c/launches_r01c.csv: `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum` of this same
    workload, eager launches).  Returns (bytes, n_launches) or (None, 0)."""
    import csv
    path = os.path.join(ROOT, "profiles", "r01c", "launches_r01c.csv")
    if not os.path.exists(path):
        return None, 0
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    if not rows:
        return None, 0
    hdr = rows[0]
    ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    ii = hdr.index("ID")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    launches = {}
    for r in rows[1:]:
        d = launches.setdefault(int(r[ii]), {"name": r[ik]})
        if r[im].startswith("dram__bytes"):
            d[r[im]] = float(r[iv].replace(",", "")) * scale.get(r[iu], 1.0)
    ids = sorted(launches)
    stems = [i for i in ids if "stem_fused" in launches[i]["name"]]
    if len(stems) < 2:
        return None, 0
    a, b = stems[-2], stems[-1]                       # one complete forward between two stem launches
    convs = [launches[i] for i in ids if a <= i < b and ("conv_tc_kernel" in launches[i]["name"])]
    total = sum(c.get("dram__bytes_read.sum", 0.0) + c.get("dram__bytes_write.sum", 0.0) for c in convs)
    return total, len(convs)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--profile-layers", default="", help="write a per-launch timing table (JSON) to this path")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200): the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", init_method="env://", device_id=dev)

    model = build_model(dev)
    B = args.batch
    gen = torch.Generator().manual_seed(100 + rank)
    host_u8 = torch.randint(0, 256, (B, 3, SIZE, SIZE), dtype=torch.uint8, generator=gen).pin_memory()
    x_dev = (host_u8.to(dev).float() / 256.0).contiguous()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        for _ in range(args.warmup):
            io, p, _ = model(x_dev)
        torch.cuda.synchronize()
        plan = model.engine().plan_for(x_dev)
        launches = plan.launches_per_forward()

        # ---------------- device-resident throughput ("value") ----------------
        sampler = ClockSampler(local)
        barrier()
        sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            io, p, _ = model(x_dev)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        clocks = sampler.stop()
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms_step = float(ms.item()) / args.steps
        value = world * B / (ms_step / 1e3)

        # ---------------- end to end through the public API with host buffers ----------------
        no = io.shape[-1]
        top_host = torch.empty((B, no), dtype=torch.float32).pin_memory()

        # double-buffered input: the H2D copy of batch i+1 runs on a copy stream while batch i computes; every
        # timed step issues exactly one H2D copy (pinned uint8 host -> device) and one D2H read of its result
        copy_stream = torch.cuda.Stream()
        dev_u8 = [torch.empty_like(host_u8, device=dev) for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]      # H2D into buffer j finished
        consumed = [torch.cuda.Event() for _ in range(2)]   # buffer j has been converted (may be overwritten)
        main = torch.cuda.current_stream()

        def issue_copy(j):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[j])
                dev_u8[j].copy_(host_u8, non_blocking=True)
                ready[j].record(copy_stream)

        for j in range(2):
            consumed[j].record(main)
        issue_copy(0)
        state = {"i": 0}

        def e2e_step():
            j = state["i"] & 1
            main.wait_event(ready[j])
            # uint8 batch straight into the model: the stem kernel applies the reference's "/ 256.0"
            # (train.py:348 / test.py:95) while it builds its im2col tile, no fp32 image is materialised
            out, _, _ = model(dev_u8[j])
            consumed[j].record(main)
            issue_copy(j ^ 1)                                             # prefetch the next batch
            idx = out[..., 4].argmax(dim=1)                               # best-objectness row per image
            top = out[torch.arange(B, device=dev), idx]
            top_host.copy_(top, non_blocking=True)
            state["i"] += 1

        for _ in range(3):
            e2e_step()
        barrier()
        e0.record()
        for _ in range(args.steps):
            e2e_step()
        e1.record()
        barrier()
        ms2 = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
        e2e_value = world * B / (float(ms2.item()) / args.steps / 1e3)

        # ---------------- roofline of the dominant kernel family (tcgen05 convs) ----------------
        conv_ms, conv_flops, n_convs = plan.time_tc_convs(x_dev, iters=max(3, min(10, args.steps)))
        pk = peaks()
        achieved = conv_flops / (conv_ms / 1e3) / 1e12
        roofline = {"bound": "tensor", "achieved": achieved, "peak": pk["tflops"], "unit": "TFLOP/s",
                    "frac": achieved / pk["tflops"], "traffic": None, "peak_source": pk["source"] + " (sustained bf16)",
                    "traffic_unit": "bytes per step, DRAM read+write of the conv launches (ncu, profiles/r01c)",
                    "kernel": "conv_tc_kernel (all %d tcgen05 conv launches of one step, back to back)" % n_convs,
                    "kernel_ms_per_step": conv_ms, "share_of_step": conv_ms / ms_step,
                    "algorithmic_flops_per_step": conv_flops}

        tb, tn = ncu_conv_traffic()
        if tb is not None and tn == n_convs and B == 32:
            roofline["traffic"] = tb
        if args.profile_layers and rank == 0:
            rows = plan.profile_layers(x_dev, reps=10)
            os.makedirs(os.path.dirname(os.path.abspath(args.profile_layers)), exist_ok=True)
            with open(args.profile_layers, "w") as f:
                json.dump({"batch": B, "size": SIZE, "ms_per_step": ms_step, "rows": rows}, f, indent=1)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu, _ = cpu_reference_rate(max_seconds=15.0)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        flops_img = conv_flops / B
        line = {
            "metric": "images/sec (640x640) yolov3 inference", "value": value, "unit": "images/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": "yolov3.cfg Darknet-53 inference bs=%d 640x640 per GPU, fp16 fused conv path "
                                   "(BASELINE configs[1])" % B,
                       "global_batch": B * world, "parallelism": "replicas x%d" % world,
                       "l2": "inputs (157 MB fp32 batch) and activations (>4 GB/step) exceed the 126 MB L2",
                       "cuda_graph": bool(getattr(model, "use_cuda_graph", True)),
                       "tcgen05_conv_flops_per_image": flops_img,
                       "conv_roofline_frac_of_step": (world * 0 + value / world) * FLOPS_PER_IMAGE / 1e12 / pk["tflops"]},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": int(host_u8.numel()),
                    "d2h_bytes_per_step": int(top_host.numel() * 4)},
            "gpu_launches": launches * args.steps,
            "roofline": roofline,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
