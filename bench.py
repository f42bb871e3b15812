#!/usr/bin/env python
"""bench.py -- keypoints+descriptors per second of the 3DFeat-Net detect-and-describe hot path on B200.

    python bench.py --gpus N --steps K --warmup W            (ours; N>1 under torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W   (the path's CPU statement on the host cores)

Workload (BASELINE.json configs[2], "C3"): per GPU a batch of 64 synthetic Oxford-shape clouds of 16384 points;
FPS to 512 clusters, ball query r=2.0 / 64 samples, detector (attention + orientation) and 32-D descriptor forward,
eval-mode BN, seed-0 random-init weights.  A step is one pass over one batch; batches shard across ranks with no
data-path collective (weak scaling).  Prints ONE JSON line (rank 0).
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


def workload_name(B, N, M, S):
    return ("C3: %d Oxford-shape clouds/GPU, %d pts, %d clusters x %d nsample, FPS+ballquery+detector+descriptor fwd" % (B, N, M, S))


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="clouds per GPU per step")
    ap.add_argument("--points", type=int, default=16384)
    ap.add_argument("--clusters", type=int, default=512)
    ap.add_argument("--nsample", type=int, default=64)
    ap.add_argument("--precision", default=os.environ.get("F3D_PRECISION", "bf16x3"), choices=["fp32", "bf16x3"],
                    help="bf16x3 = tcgen05 with split-bf16 fp32 emulation (default); fp32 = exact CUDA-core FFMA path")
    ap.add_argument("--graph", type=int, default=1, help="replay the step from a CUDA graph")
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

    def summary(self):
        sm = sorted(s[0] for s in self.samples)
        mx = [s[1] for s in self.samples]
        reasons = sorted({r for s in self.samples for r in s[2]})
        return dict(sm_mhz=(sm[len(sm) // 2] if sm else None), sm_max_mhz=(max(mx) if mx else None), reasons=reasons,
                    samples=len(self.samples), source="nvml" if self.nvml is not None else "nvidia-smi")


def cpu_reference_pass(xyz_np, params, clusters, nsample, radius=2.0):
    """The path's CPU statement: C oracle ops (OpenMP) + PyTorch-CPU fp32 network, all host threads."""
    import torch
    from oracle import net as onet

    t0 = time.perf_counter()
    for c0 in range(0, len(xyz_np), 8):  # 8 clouds at a time: the (8,512,64,256) fp32 activations stay at 268 MB
        onet.inference_model(xyz_np[c0:c0 + 8], params, num_clusters=clusters, radius=radius, nsample=nsample)
    return time.perf_counter() - t0


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
    sample = max(1, min(args.cpu_sample, args.batch))
    xyz = synth.make_batch(sample, args.points, seed0=1000)
    params = onet.to_torch(onet.init_params(seed=0))
    for _ in range(min(args.warmup, 1)):
        cpu_reference_pass(xyz[:8], params, args.clusters, args.nsample)
    steps = max(1, min(args.steps, 5))
    t = [cpu_reference_pass(xyz, params, args.clusters, args.nsample) for _ in range(steps)]
    sec = sum(t) / len(t)
    value = sample * args.clusters / sec
    line = dict(metric="keypoints+descriptors/sec", value=value, unit="keypoints/s", n_gpus=args.gpus, steps=steps,
                warmup=min(args.warmup, 1), ms_per_step=sec * 1e3, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f32", data="synthetic", impl="reference",
                config=dict(workload=workload_name(args.batch, args.points, args.clusters, args.nsample), clouds_per_step=sample,
                            precision="fp32", parallelism="rank 0 only, all host threads"),
                cpu_baseline=dict(value=value, unit="keypoints/s", cores=max(cores, oops.num_threads()), kind="port",
                                  sample="%d clouds of %d points per step (of the %d-cloud batch)" % (sample, args.points, args.batch)),
                e2e=dict(value=value, unit="keypoints/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    emit(line)


def emit(line):
    """The ONE JSON line goes to the process's original stdout; everything else written to fd 1 meanwhile (NCCL's version
    banner, library chatter) has been diverted to stderr by main()."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    args = parse()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)  # fd 1 -> stderr for the rest of the run
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch

    dist = importlib.import_module("3dfeatnet_b200.dist")
    synth = importlib.import_module("3dfeatnet_b200.synth")
    pipe_mod = importlib.import_module("3dfeatnet_b200.pipeline")
    rank, local_rank, world = dist.init("nccl")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    B, N, M, S = args.batch, args.points, args.clusters, args.nsample

    xyz = synth.make_batch(B, N, seed0=1000 + rank * B)
    pipe = pipe_mod.DetectDescribePipeline(B, N, num_clusters=M, nsample=S, precision=args.precision, device=dev,
                                           use_graph=bool(args.graph), seed=0)
    pipe.h_xyz.copy_(torch.as_tensor(xyz))
    pipe.xyz.copy_(pipe.h_xyz)
    flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device=dev)  # 1.5 x the 126 MB L2

    def l2_flush():
        flush.fill_(1)

    pipe.step()  # builds the weight images (once per set of weights)
    pipe.step()  # counts the launches of a steady-state step
    for _ in range(max(args.warmup, 3)):
        pipe.step()
    torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()

    # ---- device-resident timing: K steps, each bracketed by events, L2 flushed between steps -------------------
    L = pipe.L
    dist.barrier()
    torch.cuda.synchronize()
    L.f3d_reset_launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for s, e in ev:
        l2_flush()
        s.record()
        pipe.step()
        e.record()
    torch.cuda.synchronize()
    dist.barrier()
    dev_ms = sum(s.elapsed_time(e) for s, e in ev)
    launches = pipe.launches_per_step * args.steps
    dev_ms = dist.max_over_ranks(dev_ms, dev)

    # ---- per-stage breakdown (eager, events between the C-ABI calls) + the dominant kernel alone ----------------------
    stage_ms = {k: 0.0 for k in pipe_mod.STAGES}
    reps = min(args.steps, 10)
    rows_ms = []
    if args.precision == "bf16x3":
        L.f3d_debug_time_detector_rows(1)
    for _ in range(reps):
        l2_flush()
        evs = []
        pipe.step(events=evs)
        torch.cuda.synchronize()
        for i, k in enumerate(pipe_mod.STAGES):
            stage_ms[k] += evs[i].elapsed_time(evs[i + 1]) / reps
        if args.precision == "bf16x3":
            rows_ms.append(float(L.f3d_debug_detector_rows_ms()))
    L.f3d_debug_time_detector_rows(0)

    # ---- end to end through the public call: pinned HOST buffers in, pinned HOST buffers out, every step; the H2D of
    # step i+1 and the D2H of step i-1 overlap the compute of step i (3 streams, double buffers); L2 flushed per step
    pipe.warm_host_graphs()
    pipe.run_host_steps(3, flush=l2_flush)
    torch.cuda.synchronize()
    dist.barrier()
    e2e_ms, h_out = pipe.run_host_steps(args.steps, flush=l2_flush)
    dist.barrier()
    e2e_ms = dist.max_over_ranks(e2e_ms, dev)
    # the host result of the overlapped loop equals the device-resident result of the same batch
    ref_rows = torch.cat([pipe.keypoints, pipe.attention[..., None], pipe.orientation[..., None], pipe.features], dim=2).cpu()
    assert torch.equal(h_out, ref_rows), "end-to-end output differs from the device-resident pass"

    sampler.stop_flag = True
    sampler.join(timeout=2)

    if rank != 0:
        dist.shutdown()
        return
    peaks = load_peaks()
    units = B * M * world * args.steps
    value = units / (dev_ms * 1e-3)
    e2e_value = units / (e2e_ms * 1e-3)
    rows = B * M * S
    if rows_ms and min(rows_ms) > 0:
        # dominant kernel: det_rows_tc_kernel (conv 3->64->128->256 + max-pool of the detector), timed alone by CUDA events
        # on its launch stream.  achieved = ALGORITHMIC flops (2 * rows * 41152) / duration; the kernel
        # executes 3 bf16 MMAs per algorithmic MAC (hi*hi + hi*lo + lo*hi), reported as executed_frac.
        k_ms = sum(rows_ms) / len(rows_ms)
        k_flops = rows * FLOPS_DET_ROW
        achieved = k_flops / (k_ms * 1e-3) / 1e12
        peak = peaks["bf16"]
        # traffic: dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the committed ncu --set full capture
        # (profiles/r01_n_kernels_ncu_summary.md, same batch): 21.56 MB + 0.20 MB; only valid for the default workload
        traffic = 21.56e6 + 0.20e6 if (B, N, M, S) == (64, 16384, 512, 64) else None
        roofline = dict(bound="tensor", kernel="det_rows_tc_kernel", achieved=achieved, peak=peak, unit="TFLOP/s", frac=achieved / peak,
                        traffic=traffic, traffic_unit="bytes of DRAM per launch (ncu)", peak_source="%s bf16 burst (MEASURED_PEAKS.json)" % peaks["source"], flops_per_launch=k_flops,
                        ms_per_launch=k_ms, executed_tensor_tflops=3 * achieved, executed_frac=3 * achieved / peak)
    else:
        det_flops = rows * FLOPS_DET_ROW + B * M * FLOPS_DET_CLUSTER
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
                            precision=args.precision, parallelism="batch-sharded dp%d, no collective" % world),
                e2e=dict(value=e2e_value, unit="keypoints/s", h2d_bytes_per_step=pipe.h2d_bytes, d2h_bytes_per_step=pipe.d2h_bytes,
                         ms_per_step=e2e_ms / args.steps,
                         how="pinned host xyz -> H2D -> pipeline -> D2H of [xyz|att|ori|desc] rows, every step; copies of "
                             "neighbouring steps overlap compute on 3 streams; includes a 192 MiB L2 flush per step"),
                gpu_launches=launches, stage_ms=stage_ms, roofline=roofline, clocks=sampler.summary())
    if args.precision != "fp32":  # the exact-fp32 (CUDA-core FFMA) path on the same batch, for reference
        pipe32 = pipe_mod.DetectDescribePipeline(B, N, num_clusters=M, nsample=S, precision="fp32", device=dev, use_graph=False, seed=0)
        pipe32.xyz.copy_(pipe.xyz)
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
        d_att = ((pipe32.attention - pipe.attention).abs() / pipe32.attention.abs().amax(dim=1, keepdim=True).clamp_min(1e-30)).max().item()
        d_feat = (pipe32.features - pipe.features).abs().max().item()
        line["fp32_path"] = dict(value=B * M / (min(t32) * 1e-3), unit="keypoints/s (this rank)", ms_per_step=min(t32),
                                 max_rel_attention_diff=d_att, max_abs_descriptor_diff=d_feat,
                                 note="exact fp32 FFMA kernels on the same batch; the bf16x3 result differs by the amounts shown")
    if not args.no_cpu_baseline:
        from oracle import net as onet, ops as oops

        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        sample = max(1, args.cpu_sample)
        cpu_params = onet.to_torch(onet.init_params(seed=0))
        sample = min(sample, B)
        cpu_xyz = xyz[:sample]
        cpu_reference_pass(cpu_xyz[:8], cpu_params, M, S)  # warm-up (thread pools, allocator)
        sec = min(cpu_reference_pass(cpu_xyz, cpu_params, M, S) for _ in range(2))
        line["cpu_baseline"] = dict(value=sample * M / sec, unit="keypoints/s", cores=max(cores, oops.num_threads()), kind="port",
                                    sample="%d of the %d clouds of one step (oracle C ops + PyTorch-CPU fp32 net)" % (sample, B))
    emit(line)
    dist.shutdown()


if __name__ == "__main__":
    main()
