"""Stage-2 training step at BASELINE.json configs[3] (C4): 6 triplets per GPU = 18 clouds x 4096 points, 512 clusters x 64
samples, attention + rotation, BN batch statistics, triplet loss, backward, ONE flat-gradient all-reduce (NCCL), TF-1 Adam.
The sampling / grouping operators, the conv+BN+ReLU layers (forward and backward), the loss and Adam are the CUDA kernels of
this repository (csrc/train_layers.cu, csrc/train.cu); pooling / concat / rotation glue is torch autograd.
The whole step (incl. the NCCL all-reduce) is replayed as ONE CUDA graph (F3D_TRAIN_GRAPH=0: eager launches).
F3D_TRAIN_UNFUSED=1 times the op-by-op torch statement instead; F3D_TRAIN_PROFILE=1 prints the top kernels of one eager step.

    python tools/train_bench.py                      (1 GPU)
    torchrun --nproc-per-node N tools/train_bench.py (N GPUs)
Prints one JSON line on rank 0 (steps/s, clouds/s; time = max over ranks, CUDA events)."""
import importlib, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
dist = importlib.import_module("3dfeatnet_b200.dist"); synth = importlib.import_module("3dfeatnet_b200.synth")
f3 = importlib.import_module("3dfeatnet_b200.models.feat3dnet")

rank, local_rank, world = dist.init("nccl")
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
B, N, M = 6, 4096, 512
net = f3.Feat3dNet({'num_clusters': M}, device=dev, seed=0).train_mode()
a, p, n = (torch.as_tensor(synth.make_batch(B, N, seed0=s + 100 * rank)).to(dev) for s in (1, 2, 3))
torch.backends.cuda.matmul.allow_tf32 = False
UNFUSED = os.environ.get("F3D_TRAIN_UNFUSED") == "1"
if UNFUSED:
    importlib.import_module("3dfeatnet_b200.models.layers").FUSED_TRAINING = False
    net.param['fused_loss'] = False


def step():
    xyz, feats, att, ep = net.get_train_model(a, p, n, True)
    loss, ep = net.get_loss(xyz, feats, att, ep)
    net.get_train_op(loss, lr=1e-5, end_points=ep, grad_hook=dist.allreduce_sum_, grad_scale=1.0 / world)
    return loss


GRAPH = os.environ.get("F3D_TRAIN_GRAPH", "1") == "1" and not UNFUSED and os.environ.get("F3D_TRAIN_PROFILE") != "1"
if GRAPH:   # the whole step (~340 launches) as one CUDA graph: removes the CPU launch gaps (eager: F3D_TRAIN_GRAPH=0)
    replay = net.capture_train_step(a, p, n, lr=1e-5, grad_hook=dist.allreduce_sum_, grad_scale=1.0 / world)
    step = replay
for _ in range(3):
    step()
torch.cuda.synchronize(); dist.barrier()
K = 10
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(K):
    loss = step()
e.record(); torch.cuda.synchronize(); dist.barrier()
ms = dist.max_over_ranks(s.elapsed_time(e), dev) / K
if rank == 0:
    print(json.dumps(dict(workload="C4 stage-2 train step, 6 triplets/GPU x 4096 pts, 512 clusters x 64", n_gpus=world, ms_per_step=ms,
                          steps_per_s=1e3 / ms, clouds_per_s=3 * B * world * 1e3 / ms, loss=float(loss.detach()),
                          mlp_backend="torch matmul/autograd layers" if UNFUSED else "csrc/train_layers.cu + train_tc.cu + train.cu, torch glue", cuda_graph=GRAPH,
                          peak_mem_gb=torch.cuda.max_memory_allocated() / 2 ** 30)))
if rank == 0 and os.environ.get("F3D_TRAIN_PROFILE") == "1" and not GRAPH:
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        step(); torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
    seq = [(e.name.split("(")[0][-40:], e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total)
           for e in prof.events() if ("lin_tc" in e.name or "wgrad" in e.name or "conv_fwd" in e.name) and "Memcpy" not in e.name]
    print("per-call us:", " ".join("%s=%.0f" % (n.replace("f3d::", "").replace("void ", ""), t) for n, t in seq))
