#!/usr/bin/env python
"""bench.py -- keypoints+descriptors per second of the 3DFeat-Net detect-and-describe hot path on B200, and the stage-2
training step beside it.

    python bench.py --gpus N --steps K --warmup W            (ours; N>1 under torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W   (the path's CPU statement on the host cores)
    python bench.py --workload train|infer|both              (default both: the inference line with a nested "train" object)

Inference workload (BASELINE.json configs[2], "C3"): per GPU a batch of 64 synthetic Oxford-shape clouds of 16384 points;
FPS to 512 clusters, ball query r=2.0 / 64 samples, detector (attention + orientation) and 32-D descriptor forward,
eval-mode BN, seed-0 random-init weights.  A step is one pass over one batch; batches shard across ranks with no
data-path collective (weak scaling).
Training workload (configs[3], "C4"): per GPU 6 triplets = 18 clouds of 4096 points, 512 clusters x 64; forward with BN batch
statistics, triplet loss, backward, ONE NCCL all-reduce of the 107 619-float gradient, TF-1 Adam -- the whole step replayed as
one CUDA graph with the all-reduce inside (reference train.py:142-158, models/feat3dnet.py:359-375).
Prints ONE JSON line (rank 0).
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FLOPS_DET_ROW, FLOPS_DET_CLUSTER = 2 * 41152, 2 * 41152
FLOPS_DESC_ROW, FLOPS_DESC_CLUSTER = 2 * 10336, 2 * (8192 + 4096)  # split-weight form
TRAIN_B, TRAIN_N, TRAIN_M = 6, 4096, 512                             # reference config.py:3, train.py:15,39
GRAD_FLOATS = 107619


def workload_name(B, N, M, S):
    return ("C3: %d Oxford-shape clouds/GPU, %d pts, %d clusters x %d nsample, FPS+ballquery+detector+descriptor fwd" % (B, N, M, S))


TRAIN_WORKLOAD = ("C4: stage-2 training step, %d triplets/GPU = %d clouds x %d pts, %d clusters x 64 nsample; fwd (BN batch stats) + "
                  "triplet loss + bwd + NCCL all-reduce of %d fp32 gradients + TF-1 Adam" % (TRAIN_B, 3 * TRAIN_B, TRAIN_N, TRAIN_M, GRAD_FLOATS))


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="both", choices=["both", "infer", "train"],
                    help="both (default): the inference line (the headline metric) with the training step nested under \"train\"")
    ap.add_argument("--batch", type=int, default=64, help="clouds per GPU per step")
    ap.add_argument("--points", type=int, default=16384)
    ap.add_argument("--clusters", type=int, default=512)
    ap.add_argument("--nsample", type=int, default=64)
    ap.add_argument("--precision", default=os.environ.get("F3D_PRECISION", "bf16x3"), choices=["fp32", "bf16x3"],
                    help="bf16x3 = tcgen05 with split-bf16 fp32 emulation (default); fp32 = exact CUDA-core FFMA path")
    ap.add_argument("--graph", type=int, default=1, help="replay the step from a CUDA graph")
    ap.add_argument("--pipelined", type=int, default=1,
                    help="1 (default): software-pipelined step -- FPS of batch i+1 runs beside the contractions of batch i; 0: serial step")
    ap.add_argument("--cpu-sample", type=int, default=64,
                    help="clouds in the bounded CPU-baseline sample (default: the 64 clouds of one step, ~4 s per pass on 16 cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d.get("hbm_gbs", 6650.0), bf16=d.get("bf16_tflops", 1590.0),
                    bf16_sustained=d.get("bf16_tflops_sustained", 1400.0), source="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback")


def profiled_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, from the committed `ncu --set full` capture of this same
    command (profiles/ncu_traffic.json, written by tools/ncu_summary.py next to the summary it was read from); None if absent."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None, None
    d = json.load(open(p))
    k = d.get("kernels", {}).get(kernel)
    return (k["dram_bytes"], d.get("capture")) if k else (None, d.get("capture"))


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons sampled DURING the timed region (B200_PROFILING.md): NVML polled every ~2 ms (the timed
    region lasts tens of milliseconds, too short for nvidia-smi subprocesses), nvidia-smi as a fallback."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.samples, self.stop_flag = gpu_index, [], False
        self.nvml = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nvml = None

    def _reasons(self):
        n = self.nvml
        try:
            r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
        except Exception:
            r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        out = []
        for name, bit in (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20)):
            if r & bit:
                out.append(name)
        return out

    def run(self):
        if self.nvml is not None:
            while not self.stop_flag:
                try:
                    self.samples.append((self.nvml.nvmlDeviceGetClockInfo(self.handle, self.nvml.NVML_CLOCK_SM), self.max_sm, self._reasons()))
                except Exception:
                    pass
                time.sleep(0.002)
            return
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                o = [x.strip() for x in o]
                self.samples.append((int(o[0]), int(o[1]), [n for i, n in enumerate(names) if o[2 + i].lower().startswith("active")]))
            except Exception:
                pass
            time.sleep(0.05)

    def finish(self):
        self.stop_flag = True
        self.join(timeout=2)
        return self.summary()

    def summary(self):
        sm = sorted(s[0] for s in self.samples)
        mx = [s[1] for s in self.samples]
        reasons = sorted({r for s in self.samples for r in s[2]})
        return dict(sm_mhz=(sm[len(sm) // 2] if sm else None), sm_max_mhz=(max(mx) if mx else None), reasons=reasons,
                    samples=len(self.samples), source="nvml" if self.nvml is not None else "nvidia-smi")


# ------------------------------------------------------------------------------------------------ CPU statements (oracle)
def cpu_reference_pass(xyz_np, params, clusters, nsample, radius=2.0):
    """The path's CPU statement: C oracle ops (OpenMP) + PyTorch-CPU fp32 network, all host threads."""
    from oracle import net as onet

    t0 = time.perf_counter()
    for c0 in range(0, len(xyz_np), 8):  # 8 clouds at a time: the (8,512,64,256) fp32 activations stay at 268 MB
        onet.inference_model(xyz_np[c0:c0 + 8], params, num_clusters=clusters, radius=radius, nsample=nsample)
    return time.perf_counter() - t0


def cpu_train_pass(triplets, seed=0):
    """One stage-2 training step of the oracle (oracle/net.py: forward with BN batch statistics, loss, autograd backward, TF-1 Adam)
    on the full per-GPU batch, fp32, all host threads."""
    import torch
    from oracle import net as onet

    P = onet.to_torch(onet.init_params(seed=seed), requires_grad=True)
    t0 = time.perf_counter()
    loss, _, _ = onet.train_step(triplets[0], triplets[1], triplets[2], P, {}, num_clusters=TRAIN_M, lr=1e-5)
    return time.perf_counter() - t0, float(loss)


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The reference registers GPU-only kernels
    (tf_sampling.cpp:92, tf_grouping.cpp:125) and TF 1.15 is not installable here, so this is the oracle port
    (kind "port"): oracle/ops_oracle.c + oracle/net.py on all host cores, each step a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    import torch
    from oracle import net as onet, ops as oops

    synth = importlib.import_module("3dfeatnet_b200.synth")
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    # the CPU arm is slow (2.3 s per 64-cloud inference step, ~5 s per training step): at most 5 timed steps and 1 warm-up,
    # whatever --steps / --warmup ask for; the line says what was run
    steps = max(1, min(args.steps, 5))
    warm = min(args.warmup, 1)
    if args.workload == "train":
        trip = [synth.make_batch(TRAIN_B, TRAIN_N, seed0=s) for s in (1, 2, 3)]
        for _ in range(warm):
            cpu_train_pass(trip)
        t = [cpu_train_pass(trip)[0] for _ in range(min(steps, 3))]
        sec = sum(t) / len(t)
        value = 3 * TRAIN_B / sec
        line = dict(metric="training clouds/sec", value=value, unit="clouds/s", n_gpus=args.gpus, steps=len(t), warmup=warm,
                    ms_per_step=sec * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                    impl="reference", config=dict(workload=TRAIN_WORKLOAD, parallelism="rank 0 only, all host threads, no collective",
                                                  steps_note="CPU arm capped at 3 timed steps / 1 warm-up"),
                    cpu_baseline=dict(value=value, unit="clouds/s", cores=max(cores, oops.num_threads()), kind="port",
                                      sample="the full per-GPU batch (%d triplets) per step, oracle/net.py train_step" % TRAIN_B),
                    e2e=dict(value=value, unit="clouds/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        return emit(line)
    sample = max(1, min(args.cpu_sample, args.batch))
    xyz = synth.make_batch(sample, args.points, seed0=1000)
    params = onet.to_torch(onet.init_params(seed=0))
    for _ in range(warm):
        cpu_reference_pass(xyz[:8], params, args.clusters, args.nsample)
    t = [cpu_reference_pass(xyz, params, args.clusters, args.nsample) for _ in range(steps)]
    sec = sum(t) / len(t)
    value = sample * args.clusters / sec
    line = dict(metric="keypoints+descriptors/sec", value=value, unit="keypoints/s", n_gpus=args.gpus, steps=steps,
                warmup=warm, ms_per_step=sec * 1e3, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f32", data="synthetic", impl="reference",
                config=dict(workload=workload_name(args.batch, args.points, args.clusters, args.nsample), clouds_per_step=sample,
                            precision="fp32", parallelism="rank 0 only, all host threads",
                            steps_note="CPU arm capped at 5 timed steps / 1 warm-up (2.3 s per step)"),
                cpu_baseline=dict(value=value, unit="keypoints/s", cores=max(cores, oops.num_threads()), kind="port",
                                  sample="%d clouds of %d points per step (of the %d-cloud batch)" % (sample, args.points, args.batch)),
                e2e=dict(value=value, unit="keypoints/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    emit(line)


def emit(line):
    """The ONE JSON line goes to the process's original stdout; everything else written to fd 1 meanwhile (NCCL's version
    banner, library chatter) has been diverted to stderr by main()."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def kernel_table(timings, peaks):
    """[(name, ms, units)] of f3d_debug_kernel_timings -> per-kernel-class totals; units are algorithmic flops for the tensor kernels
    (names ending in _tc_kernel / post_tc) and algorithmic bytes otherwise."""
    agg = {}
    for name, ms, units in timings:
        if ms <= 0:
            continue
        a = agg.setdefault(name, dict(launches=0, ms=0.0, units=0.0))
        a["launches"] += 1
        a["ms"] += ms
        a["units"] += units
    rows = []
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
        tensor = name.startswith(("det_rows_tc", "desc_rows_tc", "post_tc"))
        rate = a["units"] / (a["ms"] * 1e-3)
        row = dict(kernel=name, launches=a["launches"], ms=a["ms"], bound="tensor" if tensor else "hbm",
                   achieved=rate / (1e12 if tensor else 1e9), unit="TFLOP/s" if tensor else "GB/s",
                   frac=rate / ((peaks["bf16"] * 1e12) if tensor else (peaks["hbm"] * 1e9)))
        lim = next((v for k, v in LIMITERS.items() if name.startswith(k)), None)
        if lim:  # what the committed ncu capture shows actually limits the kernel (the nominal roofline above is only the SURVEY 8d figure)
            row["limiter"] = lim
        rows.append(row)
    return rows


# the resource that limits each inference kernel according to profiles/r02_be_kernels_ncu_summary.md (ncu --set full, C3)
LIMITERS = {
    "fps_group_kernel": "serial latency: 511 dependent rounds per cloud, 0.53 us per round (one barrier + two redux trees); issue slots 21 %, DRAM 0.6 %",
    "bq_grid_query_kernel": "instruction issue 58 % (IPC 2.3) + L2 gather latency (long_scoreboard); DRAM 3.5 %, L2 hit 76 %",
    "bq_grid_build_kernel": "shared-memory counting sort (mio / lg throttle); DRAM 7 %",
    "det_rows_tc_kernel": "tensor pipe 81 % active; 3 MMAs per algorithmic MAC (bf16x3), DRAM traffic = algorithmic bytes",
    "desc_rows_tc_kernel": "instruction streams of its 22 warps (producers ~200 instructions per 64-sample tile): tensor pipe 59 %, issue slots 53 %, L1/shared 69 %",
    "post_tc_kernel": "latency: weight staging + three dependent MMA groups per 64-cluster tile; every unit below 30 %",
}


# ------------------------------------------------------------------------------------------------ inference workload (W1)
def bench_infer(args, dist, dev, rank, local_rank, world, peaks):
    import numpy as np
    import torch

    synth = importlib.import_module("3dfeatnet_b200.synth")
    pipe_mod = importlib.import_module("3dfeatnet_b200.pipeline")
    _lib = importlib.import_module("3dfeatnet_b200._lib")
    B, N, M, S = args.batch, args.points, args.clusters, args.nsample

    xyz = synth.make_batch(B, N, seed0=1000 + rank * B)
    pipe = pipe_mod.DetectDescribePipeline(B, N, num_clusters=M, nsample=S, precision=args.precision, device=dev,
                                           use_graph=bool(args.graph), seed=0)
    pipelined = bool(args.pipelined) and hasattr(pipe, "step_pipelined")
    pipe.h_xyz.copy_(torch.as_tensor(xyz))
    pipe.xyz.copy_(pipe.h_xyz)
    flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device=dev)  # 1.5 x the 126 MB L2

    def l2_flush():
        flush.fill_(1)

    pipe.step()  # builds the weight images (once per set of weights)
    pipe.step()  # counts the launches of a steady-state step
    serial_out = {k: getattr(pipe, k).clone() for k in ("keypoints", "attention", "orientation", "features", "idx", "fps_idx")}
    one_step = pipe.step
    partition = None
    if pipelined:
        # [FPS + grid build of batch i+1] beside [ball query + detector + descriptor of batch i]; the SM partition between the two is
        # measured once (a few candidates, 8 steps each) like any launch-configuration autotune, outside the timed region
        tuned = pipe.tune_pipelined()
        partition = dict(fps_ctas=pipe.fps_ctas, det_sm_limit=pipe.det_sm_limit, desc_sm_limit=pipe.desc_sm_limit,
                         candidates_ms={"%d/%d/%d" % c[:3]: round(c[3], 4) for c in tuned})
        pipe.prime_pipelined()          # FPS of the first batch (prologue, outside the timed region)
        one_step = pipe.step_pipelined
    for _ in range(max(args.warmup, 3)):
        one_step()
    torch.cuda.synchronize()
    if pipelined:  # the pipelined step computes exactly what the serial step computes
        for k, v in serial_out.items():
            assert torch.equal(getattr(pipe, k), v), "pipelined step differs from the serial step in %s" % k

    sampler = ClockSampler(local_rank)
    sampler.start()

    # ---- device-resident timing: K steps, each bracketed by events, L2 flushed between steps -------------------
    L = pipe.L
    dist.barrier()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for s, e in ev:
        l2_flush()
        s.record()
        one_step()
        e.record()
    torch.cuda.synchronize()
    dist.barrier()
    dev_ms = sum(s.elapsed_time(e) for s, e in ev)
    launches = pipe.launches_per_step * args.steps
    dev_ms = dist.max_over_ranks(dev_ms, dev)

    # ---- per-stage breakdown (eager serial step, events between the C-ABI calls) + every hot kernel alone (CUDA events on its
    # own launch stream, f3d_debug_kernel_timer) -----------------------------------------------------------------------------
    stage_ms = {k: 0.0 for k in pipe_mod.STAGES}
    reps = min(args.steps, 10)
    L.f3d_debug_kernel_timer(1)
    for _ in range(reps):
        l2_flush()
        evs = []
        pipe.step(events=evs)
        torch.cuda.synchronize()
        for i, k in enumerate(pipe_mod.STAGES):
            stage_ms[k] += evs[i].elapsed_time(evs[i + 1]) / reps
    timings = _lib.kernel_timings()
    L.f3d_debug_kernel_timer(0)
    kernels = kernel_table(timings, peaks)
    for k in kernels:  # per launch
        k["ms"] /= max(1, k["launches"])
        k["launches"] = 1

    # ---- end to end through the public call: pinned HOST buffers in, pinned HOST buffers out, every step; the H2D of later
    # steps and the D2H of earlier ones overlap the compute of step i (3 streams, ring buffers); L2 flushed per step
    pipe.host_pipelined = pipelined
    e2e_flush, e2e_l2 = l2_flush, "includes a 192 MiB L2 flush per step"
    if pipelined:
        # every step's input arrives from the host into the next slot of a ring of device buffers that is larger than the L2 (a slot is
        # rewritten after `ring` steps), so no step finds its input or a previous step's working set in the cache and no flush is needed
        ring = -(-int(1.2 * 126e6) // (B * N * 12))
        pipe.host_ring = max(4, ring + (ring & 1))
        e2e_flush = None
        e2e_l2 = "inputs land in a ring of %d device buffers = %.0f MB > the 126 MB L2, no flush" % (pipe.host_ring, pipe.host_ring * B * N * 12 / 1e6)
    pipe.warm_host_graphs()
    pipe.run_host_steps(3, flush=e2e_flush)
    torch.cuda.synchronize()
    dist.barrier()
    e2e_ms, h_out = pipe.run_host_steps(args.steps, flush=e2e_flush)
    dist.barrier()
    e2e_ms = dist.max_over_ranks(e2e_ms, dev)
    # the host result of the overlapped loop equals the device-resident result of the same batch
    ref_rows = torch.cat([serial_out["keypoints"], serial_out["attention"][..., None], serial_out["orientation"][..., None],
                          serial_out["features"]], dim=2).cpu()
    assert torch.equal(h_out, ref_rows), "end-to-end output differs from the device-resident pass"

    # the same host<->device traffic with NO compute, all ranks at once: what the box's host links deliver to this rank when every GPU
    # copies (a step cannot be shorter than this; at N = 8 this, not the kernels, is what bounds the end-to-end number)
    hp = pipe._host_pipe()
    dbuf = torch.empty_like(pipe.xyz)
    hbuf = pipe.h_xyz
    dout, hout = hp["d_out"][0], hp["h_out"][0]
    sc, ec = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for timed in (False, True):
        torch.cuda.synchronize()
        dist.barrier()
        with torch.cuda.stream(hp["s_h2d"]):
            sc.record()
            for _ in range(args.steps):
                dbuf.copy_(hbuf, non_blocking=True)
        with torch.cuda.stream(hp["s_d2h"]):
            for _ in range(args.steps):
                hout.copy_(dout, non_blocking=True)
            hp["s_d2h"].wait_stream(hp["s_h2d"])
            ec.record()
        torch.cuda.synchronize()
    copies_ms = dist.max_over_ranks(sc.elapsed_time(ec), dev) / args.steps
    dist.barrier()

    clocks = sampler.finish()
    if rank != 0:
        return None
    units = B * M * world * args.steps
    value = units / (dev_ms * 1e-3)
    e2e_value = units / (e2e_ms * 1e-3)
    top = next((k for k in kernels if k["kernel"] == "det_rows_tc_kernel"), None)
    if top is not None:
        # dominant kernel: det_rows_tc_kernel (conv 3->64->128->256 + max-pool of the detector).  achieved = ALGORITHMIC flops
        # (2 * rows * 41152) / duration; it executes 3 bf16 MMAs per algorithmic MAC (hi*hi + hi*lo + lo*hi).
        traffic, capture = profiled_traffic("det_rows_tc_kernel") if (B, N, M, S) == (64, 16384, 512, 64) else (None, None)
        roofline = dict(bound="tensor", kernel="det_rows_tc_kernel", achieved=top["achieved"], peak=peaks["bf16"], unit="TFLOP/s",
                        frac=top["frac"], traffic=traffic, traffic_unit="bytes of DRAM per launch", traffic_source=capture,
                        peak_source="%s bf16 burst (MEASURED_PEAKS.json)" % peaks["source"],
                        flops_per_launch=B * M * S * FLOPS_DET_ROW, ms_per_launch=top["ms"],
                        executed_tensor_tflops=3 * top["achieved"], executed_frac=3 * top["frac"])
    else:
        det_flops = B * M * S * FLOPS_DET_ROW + B * M * FLOPS_DET_CLUSTER
        det_ms = stage_ms["detector"]
        achieved = det_flops / (det_ms * 1e-3) / 1e12
        roofline = dict(bound="tensor", kernel="f3d_detector_forward (fp32 FFMA path: det_rows_fp32 + det_post_fp32)", achieved=achieved,
                        peak=peaks["bf16"], unit="TFLOP/s", frac=achieved / peaks["bf16"], traffic=None,
                        peak_source="%s bf16 burst (MEASURED_PEAKS.json); this path runs on the fp32 FFMA pipe" % peaks["source"],
                        flops_per_launch=det_flops, ms_per_launch=det_ms)
    line = dict(metric="keypoints+descriptors/sec", value=value, unit="keypoints/s", n_gpus=world, steps=args.steps,
                warmup=max(args.warmup, 3), ms_per_step=dev_ms / args.steps, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f32" if args.precision == "fp32" else "f32 via bf16x3 tensor-core split (fp32 accumulate)",
                data="synthetic",
                config=dict(workload=workload_name(B, N, M, S), l2="flushed between timed steps (192 MiB write)", cuda_graph=bool(args.graph),
                            precision=args.precision, parallelism="batch-sharded dp%d, no collective" % world, sm_partition=partition,
                            step=("software-pipelined: each step = [FPS + grid build of batch i+1] beside [ball query + detector + descriptor "
                                  "of batch i]; K steps do K of each, the first batch's FPS is a prologue outside the timed region; outputs "
                                  "bit-identical to the serial step" if pipelined else "serial: FPS -> ball query -> detector -> descriptor")),
                e2e=dict(value=e2e_value, unit="keypoints/s", h2d_bytes_per_step=pipe.h2d_bytes, d2h_bytes_per_step=pipe.d2h_bytes,
                         ms_per_step=e2e_ms / args.steps,
                         copies_alone_ms_per_step=copies_ms,
                         copies_alone_GBps_per_gpu=(pipe.h2d_bytes + pipe.d2h_bytes) / (copies_ms * 1e-3) / 1e9,
                         copies_alone_note="the step's H2D + D2H with no compute, every rank copying at once (max over ranks): the floor the "
                                           "box's host links put under an end-to-end step at this N",
                         how="pinned host xyz -> H2D -> pipeline -> D2H of [xyz|att|ori|desc] rows, every step; copies of "
                             "neighbouring steps overlap compute on 3 streams; " + e2e_l2),
                gpu_launches=launches, stage_ms=stage_ms, roofline=roofline, kernels=kernels, clocks=clocks)
    if args.precision != "fp32":  # the exact-fp32 (CUDA-core FFMA) path on the same batch, for reference
        pipe32 = pipe_mod.DetectDescribePipeline(B, N, num_clusters=M, nsample=S, precision="fp32", device=dev, use_graph=False, seed=0)
        pipe32.xyz.copy_(pipe.h_xyz)
        for _ in range(2):
            pipe32.step()
        torch.cuda.synchronize()
        t32 = []
        for _ in range(3):
            l2_flush()
            s0, e0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            pipe32.step()
            e0.record()
            torch.cuda.synchronize()
            t32.append(s0.elapsed_time(e0))
        line["fp32_path"] = dict(value=B * M / (min(t32) * 1e-3), unit="keypoints/s (this rank)", ms_per_step=min(t32),
                                 note="exact fp32 FFMA kernels on the same batch (both paths are held to the oracle by `parity` below)")
        del pipe32
    if not args.no_cpu_baseline:
        from oracle import net as onet, ops as oops, parity

        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        # ---- parity of THIS run's output against the fp64 oracle, under the one tolerance table (oracle/parity.py): asserted
        npar = min(16, B)
        P64 = onet.to_torch(onet.init_params(seed=0), torch.float64)
        ref = parity.oracle_forward(onet, xyz[:npar], P64, M, nsample=S)
        out = dict(attention=serial_out["attention"][:npar], orientation=serial_out["orientation"][:npar], features=serial_out["features"][:npar])
        assert np.array_equal(serial_out["fps_idx"][:npar].cpu().numpy(), ref["fps_idx"]), "FPS indices differ from the oracle"
        assert np.array_equal(serial_out["idx"][:npar].cpu().numpy(), ref["idx"]), "ball-query indices differ from the oracle"
        own = parity.oracle_descriptor_at(onet, xyz[:npar], P64, ref["xyz"], out["orientation"], nsample=S)
        err = parity.check(parity.errors(out, ref, own), args.precision, "bench")
        line["parity"] = dict(vs="fp64 oracle (oracle/net.py) on the first %d clouds of the timed batch" % npar, indices="FPS and ball query bit-exact",
                              errors=err, tolerance=parity.TOL[args.precision], ok=True)
    if not args.no_cpu_baseline and world == 1:  # the CPU statement timed on this box's cores (rank 0 at N = 1 only)
        from oracle import net as onet, ops as oops

        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        sample = min(max(1, args.cpu_sample), B)
        cpu_params = onet.to_torch(onet.init_params(seed=0))
        cpu_xyz = xyz[:sample]
        cpu_reference_pass(cpu_xyz[:8], cpu_params, M, S)  # warm-up (thread pools, allocator)
        sec = min(cpu_reference_pass(cpu_xyz, cpu_params, M, S) for _ in range(2))
        line["cpu_baseline"] = dict(value=sample * M / sec, unit="keypoints/s", cores=max(cores, oops.num_threads()), kind="port",
                                    sample="%d of the %d clouds of one step (oracle C ops + PyTorch-CPU fp32 net)" % (sample, B))
    return line


# ------------------------------------------------------------------------------------------------ training workload (W3)
def bench_train(args, dist, dev, rank, local_rank, world, peaks):
    import torch

    synth = importlib.import_module("3dfeatnet_b200.synth")
    f3 = importlib.import_module("3dfeatnet_b200.models.feat3dnet")
    _lib = importlib.import_module("3dfeatnet_b200._lib")
    L = _lib.lib()
    torch.backends.cuda.matmul.allow_tf32 = False
    B, N, M = TRAIN_B, TRAIN_N, TRAIN_M
    net = f3.Feat3dNet({'num_clusters': M}, device=dev, seed=0).train_mode()
    trip_np = [synth.make_batch(B, N, seed0=s + 100 * rank) for s in (1, 2, 3)]
    host = [torch.as_tensor(t).pin_memory() for t in trip_np]
    a, p, n = (t.to(dev) for t in host)
    scale = 1.0 / world

    # per-kernel table of ONE eager step (CUDA events around every hot launch); it is a real step like the others
    def eager_step():
        xyz, feats, att, ep = net.get_train_model(a, p, n, True)
        loss, ep = net.get_loss(xyz, feats, att, ep)
        net.get_train_op(loss, lr=1e-5, end_points=ep, grad_hook=dist.allreduce_sum_, grad_scale=scale)
        return loss

    eager_step()
    torch.cuda.synchronize()
    L.f3d_reset_launch_count()
    L.f3d_debug_kernel_timer(1)
    eager_step()
    timings = _lib.kernel_timings()
    L.f3d_debug_kernel_timer(0)
    own_launches = int(L.f3d_launch_count())
    kernels = kernel_table(timings, peaks)

    def time_replays(replay):
        for _ in range(max(args.warmup, 3)):
            replay()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(args.steps):
            loss = replay()
        e.record()
        torch.cuda.synchronize()
        dist.barrier()
        return dist.max_over_ranks(s.elapsed_time(e), dev) / args.steps, loss

    # the serial step (sampling -> forward -> loss -> backward -> all-reduce -> Adam, one graph) on a model of its own, for the record;
    # the timed step is the software-pipelined one unless --pipelined 0: the clusters of the NEXT batch (FPS + ball query: 18 CTAs, serial
    # rounds) are computed on a side stream beside the backward pass of the current step; K steps do K samplings and K optimiser steps,
    # the first batch's sampling is a prologue outside the timed region; same bits as the serial step (tests/test_train_gpu.py)
    pipelined = bool(args.pipelined)
    sampler = ClockSampler(local_rank)
    sampler.start()
    serial_ms = None
    if pipelined:
        net_s = f3.Feat3dNet({'num_clusters': M}, device=dev, seed=0).train_mode()
        serial_ms, _ = time_replays(net_s.capture_train_step(a, p, n, lr=1e-5, grad_hook=dist.allreduce_sum_, grad_scale=scale, warmup=1))
        del net_s
        torch.cuda.empty_cache()
    replay = net.capture_train_step(a, p, n, lr=1e-5, grad_hook=dist.allreduce_sum_, grad_scale=scale, warmup=1, pipelined=pipelined)
    ms, loss = time_replays(replay)

    # end to end: the triplet batch comes from pinned host memory every step and the loss goes back to the host
    h_loss = torch.empty(1, dtype=torch.float32).pin_memory()
    s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    torch.cuda.synchronize()
    s2.record()
    for _ in range(args.steps):
        loss = replay(*host)  # H2D of the pinned triplets into the graph's input buffers (pipelined: the batch of the NEXT step), then the step
        h_loss.copy_(loss.reshape(1), non_blocking=True)
    e2.record()
    torch.cuda.synchronize()
    dist.barrier()
    e2e_ms = dist.max_over_ranks(s2.elapsed_time(e2), dev) / args.steps
    clocks = sampler.finish()

    # the collective alone (same buffer size, same communicator), to name its share of the step
    ar_us = None
    if world > 1:
        buf = torch.zeros(GRAD_FLOATS, dtype=torch.float32, device=dev)
        for _ in range(5):
            dist.allreduce_sum_(buf)
        torch.cuda.synchronize()
        dist.barrier()
        s3, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s3.record()
        for _ in range(50):
            dist.allreduce_sum_(buf)
        e3.record()
        torch.cuda.synchronize()
        ar_us = dist.max_over_ranks(s3.elapsed_time(e3), dev) / 50 * 1e3
    if rank != 0:
        return None
    clouds = 3 * B * world
    top = kernels[0] if kernels else None
    hbm_ms = sum(k["ms"] for k in kernels if k["bound"] == "hbm")
    hbm_bytes = sum(k["ms"] * 1e-3 * k["achieved"] * 1e9 for k in kernels if k["bound"] == "hbm")
    train = dict(metric="training clouds/sec", value=clouds * 1e3 / ms, unit="clouds/s", n_gpus=world, steps=args.steps,
                 warmup=max(args.warmup, 3), ms_per_step=ms, steps_per_s=1e3 / ms, higher_is_better=True, scaling="weak",
                 dtype="f32 (forward contractions: 3-way bf16 split on tcgen05; dgrad / wgrad: 2-way split; everything else fp32)",
                 data="synthetic", loss=float(loss),
                 config=dict(workload=TRAIN_WORKLOAD, cuda_graph=True, l2="per-step working set (>2 GB of activations) exceeds the 126 MB L2",
                             step=("software-pipelined: FPS + ball query of batch i+1 on a side stream beside the backward pass of batch i; K steps do K "
                                   "samplings and K optimiser steps, the first batch's sampling is a prologue outside the timed region; same bits "
                                   "as the serial step" if pipelined else "serial: sampling -> forward -> loss -> backward -> all-reduce -> Adam"),
                             serial_ms_per_step=serial_ms,
                             parallelism="data-parallel dp%d: triplets sharded, BN statistics per GPU, one SUM all-reduce + 1/world in Adam" % world),
                 e2e=dict(value=clouds * 1e3 / e2e_ms, unit="clouds/s", ms_per_step=e2e_ms, h2d_bytes_per_step=sum(t.numel() * 4 for t in host),
                          d2h_bytes_per_step=4, how="pinned host triplets -> H2D -> graph replay -> loss D2H, every step"),
                 gpu_launches=own_launches * args.steps, launches_per_step=own_launches,
                 allreduce=dict(floats=GRAD_FLOATS, inside_graph=True, alone_us=ar_us,
                                share_of_step=(ar_us * 1e-3 / ms if ar_us else 0.0),
                                note="latency-bound (430 KB): the collective is issued once per step after the last weight gradient"),
                 kernels=kernels[:12], clocks=clocks)
    if top is not None:
        train["roofline"] = dict(bound="hbm", kernel="all HBM-bound kernels of the step (lin_tc, wgrad_tc, BN passes)", achieved=hbm_bytes / (hbm_ms * 1e-3) / 1e9 if hbm_ms else None,
                                 peak=peaks["hbm"], unit="GB/s", frac=(hbm_bytes / (hbm_ms * 1e-3) / 1e9 / peaks["hbm"]) if hbm_ms else None,
                                 traffic=None, peak_source="%s HBM copy bandwidth (MEASURED_PEAKS.json)" % peaks["source"],
                                 algorithmic_bytes_per_step=hbm_bytes, ms_in_these_kernels=hbm_ms, top_kernel=top["kernel"],
                                 how="algorithmic bytes as stated at each launch site (rows x channels x 4 in and out) / CUDA-event time of the eager launches")
    if not args.no_cpu_baseline and world == 1:
        from oracle import ops as oops

        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        sec, _ = cpu_train_pass(trip_np)
        sec = min(sec, cpu_train_pass(trip_np)[0])
        train["cpu_baseline"] = dict(value=3 * B / sec, unit="clouds/s", cores=max(cores, oops.num_threads()), kind="port",
                                     sample="one full per-GPU batch (%d triplets): oracle/net.py train_step, fp32, autograd" % B)
    return train


def main():
    global _REAL_STDOUT
    args = parse()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)  # fd 1 -> stderr for the rest of the run
    if args.impl == "reference":
        return run_reference(args)

    import torch

    dist = importlib.import_module("3dfeatnet_b200.dist")
    rank, local_rank, world = dist.init("nccl")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:  # pinned host buffers of every rank on its own GPU's NUMA node (the end-to-end path moves 17 MB per step per GPU)
        dist.bind_to_gpu_numa_node(local_rank)
    peaks = load_peaks()
    line = None
    if args.workload in ("both", "infer"):
        line = bench_infer(args, dist, dev, rank, local_rank, world, peaks)
    if args.workload in ("both", "train"):
        torch.cuda.empty_cache()
        train = bench_train(args, dist, dev, rank, local_rank, world, peaks)
        if rank == 0:
            if line is None:
                line = train
                line["vs_baseline"] = None
            else:
                line["train"] = train
    if rank == 0:
        emit(line)
    dist.shutdown()


if __name__ == "__main__":
    main()
