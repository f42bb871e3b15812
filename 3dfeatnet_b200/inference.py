"""The device-side pieces of the reference's inference.py: keypoint NMS (inference.py:226-261) and the two-pass
detect -> NMS -> describe flow of compute_descriptors (inference.py:99-180) on torch CUDA tensors."""
import importlib

import torch

_ROOT = __name__.split(".")[0]
_lib = importlib.import_module("3dfeatnet_b200._lib" if _ROOT == "3dfeatnet_b200" else "_lib")
_dist = importlib.import_module("3dfeatnet_b200.dist" if _ROOT == "3dfeatnet_b200" else "dist")
_tg = importlib.import_module(("3dfeatnet_b200." if _ROOT == "3dfeatnet_b200" else "") + "tf_ops.grouping.tf_grouping")

MAX_POINTS = 30000  # inference.py:22: centres per detection pass


def nms(xyz, attention, nms_radius=0.5, min_response_ratio=1e-2, max_keypoints=1024, num_neighbors=50, return_indices=False, counts_on_device=False):
    """nms(xyz, attention) of inference.py:226-261 (args.nms_radius / min_response_ratio / max_keypoints become keywords
    with the CLI defaults, :40-47).  xyz (B,N,3), attention (B,N) CUDA float32.
    Returns (xyz_nms (B,K,3), attention_nms (B,K), num_keypoints list[int]) like the reference (+ indices on request).
    counts_on_device: num_keypoints stays a CUDA int32 tensor (B,) and the call does not wait for the device."""
    if xyz.dim() != 3 or xyz.shape[2] != 3 or attention.dim() != 2 or attention.shape != xyz.shape[:2]:
        raise ValueError("nms expects xyz (B,N,3) and attention (B,N)")
    _lib.require_cuda(xyz, attention)
    if xyz.dtype != torch.float32 or attention.dtype != torch.float32:
        raise ValueError("nms expects float32 tensors")
    xyz, attention = xyz.contiguous(), attention.contiguous()
    b, n, _ = xyz.shape
    if n < num_neighbors:
        raise ValueError("Expected n_neighbors <= n_samples (the reference's BallTree query raises too)")
    dev = xyz.device
    out_idx = torch.empty((b, max_keypoints), dtype=torch.int32, device=dev)
    out_xyz = torch.empty((b, max_keypoints, 3), dtype=torch.float32, device=dev)
    out_att = torch.empty((b, max_keypoints), dtype=torch.float32, device=dev)
    num = torch.empty((b,), dtype=torch.int32, device=dev)
    L = _lib.lib()
    ws_bytes = L.f3d_nms_workspace_bytes(b, n)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    _lib.check(L.f3d_nms(b, n, _lib.ptr(xyz), _lib.ptr(attention), float(nms_radius), float(min_response_ratio), max_keypoints,
                         num_neighbors, _lib.ptr(out_idx), _lib.ptr(out_xyz), _lib.ptr(out_att), _lib.ptr(num), _lib.ptr(ws), ws_bytes,
                         _lib.stream()), "nms")
    counts = num if counts_on_device else num.cpu().tolist()
    if return_indices:
        return out_xyz, out_att, counts, out_idx
    return out_xyz, out_att, counts


def detect_and_describe(model, point_cloud, nms_radius=0.5, min_response_ratio=1e-2, max_keypoints=1024, counts_on_device=False):
    """compute_descriptors' per-cloud body (inference.py:115-171): attention at EVERY point (centres in chunks of
    MAX_POINTS, num_clusters=-1 semantics), NMS, then descriptors at the <= max_keypoints survivors.
    point_cloud: (1,N,>=3) CUDA float32.  Returns (xyz_nms (1,K,3), features (1,K,F), attention_nms (1,K), num_keypoints);
    counts_on_device: num_keypoints stays on the device and nothing in the call waits for it (the file loop overlaps scans)."""
    xyz = point_cloud[:, :, :3].contiguous()
    n = xyz.shape[1]
    # the cloud is binned once for the ball queries of all chunks and of the descriptor pass (the model sees the same xyz tensor)
    reuse = xyz.is_cuda and hasattr(model, "packed_weights")  # (stand-in models of the CPU tests take the plain call)
    extra = {"ball_grid": _tg.BallGrid(model.param['BaseScale'], xyz, max_centres=min(n, MAX_POINTS))} if reuse else {}
    cloud = xyz if reuse else point_cloud  # the model reads columns 0..2 only; handing it xyz itself keeps the grid's tensor identity
    atts = []
    for s in range(0, n, MAX_POINTS):
        kp = xyz[:, s:s + MAX_POINTS, :].contiguous()
        _, _, att, ep = model.get_inference_model(cloud, False, keypoints=kp, fetch_features=False, **extra)
        atts.append(ep["attention"])
    attention = torch.cat(atts, dim=1)
    xyz_nms, att_nms, num = nms(xyz, attention, nms_radius, min_response_ratio, max_keypoints, **({'counts_on_device': True} if counts_on_device else {}))
    _, features, _, _ = model.get_inference_model(cloud, False, keypoints=xyz_nms, **extra)
    return xyz_nms, features, att_nms, num


# ---------------------------------------------------------------------------------------------- file formats
def load_point_cloud(path, num_cols=6):
    """data/datagenerator.py:163-182: a raw float32 file of N x num_cols values (xyz + 3 unused columns for clouds, 3 for
    keypoint files) -> (N,num_cols)."""
    import numpy as np

    a = np.fromfile(path, dtype=np.float32)
    if a.size % num_cols:
        raise ValueError("%s: %d floats is not a multiple of %d columns" % (path, a.size, num_cols))
    return a.reshape(-1, num_cols)


def save_keypoints_and_descriptors(path, xyz, features):
    """inference.py:174-177: one float32 row [x y z | descriptor] per keypoint -- the format scripts/Utils.m:56 reads."""
    import numpy as np

    xyz = xyz.detach().cpu().numpy() if torch.is_tensor(xyz) else np.asarray(xyz)
    features = features.detach().cpu().numpy() if torch.is_tensor(features) else np.asarray(features)
    rows = np.concatenate([xyz.reshape(-1, 3), features.reshape(xyz.reshape(-1, 3).shape[0], -1)], axis=1).astype(np.float32)
    rows.tofile(path)
    return rows.shape


def _load_inputs(in_path, randomize_points, seed, num_points, keypoints_path, data_dim):
    """host part of one iteration: the cloud as (1,N,data_dim) float32 (permuted / truncated as inference.py:108-116) and the fed
    keypoints (1,K,3) or None"""
    import numpy as np

    cloud = load_point_cloud(in_path, num_cols=data_dim)
    if randomize_points:  # inference.py:108-113
        cloud = cloud[np.random.default_rng(seed).permutation(cloud.shape[0])]
    if num_points > 0:  # :115-116
        cloud = cloud[:num_points]
    kp = None
    if keypoints_path is not None:
        kp = load_point_cloud(keypoints_path, num_cols=3)
        if kp.shape[0] == 0:
            raise ValueError("%s holds no keypoints" % keypoints_path)
        kp = np.ascontiguousarray(kp[None])
    return np.ascontiguousarray(cloud[None]), kp


def _describe_padded(model, cloud, kp, max_keypoints, nms_radius, min_response_ratio, device, wait=True):
    """device part: detect at every point + NMS (inference.py:118-151) or describe the fed keypoints (:153-158) -> (xyz rows, feature
    rows, number of real rows); wait=False: the number stays a device tensor when it comes from the NMS and the call only enqueues work"""
    pc = torch.as_tensor(cloud).to(device)
    if kp is None:
        xyz_nms, features, _, num = detect_and_describe(model, pc, nms_radius, min_response_ratio, max_keypoints, counts_on_device=not wait)
        return xyz_nms[0], features[0], num[0]
    xyz_nms, features, _, _ = model.get_inference_model(pc, False, keypoints=torch.as_tensor(kp).to(device))
    return xyz_nms[0], features[0], kp.shape[1]


def _describe(model, cloud, kp, max_keypoints, nms_radius, min_response_ratio, device):
    xyz, features, k = _describe_padded(model, cloud, kp, max_keypoints, nms_radius, min_response_ratio, device)
    return xyz[:k], features[:k]


def compute_descriptors_for_file(model, in_path, out_path, randomize_points=False, seed=0, max_keypoints=1024,
                                 nms_radius=0.5, min_response_ratio=1e-2, device="cuda", num_points=-1, keypoints_path=None,
                                 data_dim=6):
    """One iteration of compute_descriptors' file loop (inference.py:99-180): load .bin, (optionally) permute the points and
    keep the first `num_points` of them (:108-116), then either detect at every point + NMS (:118-151) or read the keypoints
    of `keypoints_path` (a float32 file of xyz rows, :153-158), describe, keep the num_keypoints real rows and write
    [xyz | descriptor] rows.  Returns the (rows, 3 + feature_dim) shape written."""
    cloud, kp = _load_inputs(in_path, randomize_points, seed, num_points, keypoints_path, data_dim)
    xyz, features = _describe(model, cloud, kp, max_keypoints, nms_radius, min_response_ratio, device)
    return save_keypoints_and_descriptors(out_path, xyz, features)


def compute_descriptors(model, data_dir, output_dir, data_dim=6, num_points=-1, use_keypoints_from=None, randomize_points=False,
                        nms_radius=0.5, min_response_ratio=1e-2, max_keypoints=1024, seed=0, device="cuda", rank=0, world=1):
    """compute_descriptors() of inference.py:67-180 without the TF session / argument parsing: every `<name>.bin` of
    `data_dir` -> `output_dir/<name>.bin` holding [xyz | descriptor] float32 rows.  The arguments are the CLI's (:26-58);
    with `use_keypoints_from` the keypoints of `<use_keypoints_from>/<name>_kp.bin` are described instead of detected ones.
    `model` is a Feat3dNet built the way the reference builds it for inference (num_clusters=-1, Attention=True, :81-83).
    Scans are independent, so with `world` > 1 processes (one per GPU) rank r takes the r-th contiguous slice of the sorted
    file list (dist.shard_range) and no collective is needed; the permutation seed of a file does not depend on the split.
    The host part of the loop runs beside the device part: file i+1 is read (and permuted) and the rows of file i-1 are written by two
    helper threads while the device works on file i -- same files, byte for byte, as the plain loop over compute_descriptors_for_file.
    Returns the file names this rank processed, in order (sorted; the reference takes os.listdir order)."""
    import os
    from concurrent.futures import ThreadPoolExecutor

    os.makedirs(output_dir, exist_ok=True)
    bin_files = sorted(f for f in os.listdir(data_dir) if f.endswith(".bin"))
    lo, hi = _dist.shard_range(len(bin_files), rank, world)
    mine = list(enumerate(bin_files))[lo:hi]

    def load(item):
        i, f = item
        kp_path = None if use_keypoints_from is None else os.path.join(use_keypoints_from, "%s_kp.bin" % f[:-4])
        return _load_inputs(os.path.join(data_dir, f), randomize_points, seed + i, num_points, kp_path, data_dim)

    # The device is kept one scan ahead of the host: the work of file n+1 is enqueued before the host waits for the rows of file n (their
    # D2H into a ring of pinned buffers was enqueued right behind file n's kernels), so the device never idles on file I/O or on the host.
    overlap = torch.device(device).type == "cuda" and hasattr(model, "packed_weights")  # (stand-in models of the CPU tests: plain order)
    ring, pending = [], None  # pending = (file name, pinned rows, pinned count or int, event, ring slot)

    def finish(p):
        fname, h_xyz, h_feat, h_k, ev, slot = p
        ev.synchronize()
        k = int(h_k[0]) if torch.is_tensor(h_k) else int(h_k)
        fut = writer.submit(save_keypoints_and_descriptors, os.path.join(output_dir, fname), h_xyz[:k].clone(), h_feat[:k].clone())
        written.append(fut)

    with ThreadPoolExecutor(max_workers=1) as reader, ThreadPoolExecutor(max_workers=1) as writer:
        nxt = reader.submit(load, mine[0]) if mine else None
        written = []
        for n, (i, f) in enumerate(mine):
            cloud, kp = nxt.result()
            nxt = reader.submit(load, mine[n + 1]) if n + 1 < len(mine) else None
            if not overlap:
                xyz, features = _describe(model, cloud, kp, max_keypoints, nms_radius, min_response_ratio, device)
                xyz, features = xyz.detach().cpu(), features.detach().cpu()
                written.append(writer.submit(save_keypoints_and_descriptors, os.path.join(output_dir, f), xyz, features))
                continue
            xyz, features, k = _describe_padded(model, cloud, kp, max_keypoints, nms_radius, min_response_ratio, device, wait=False)
            slot = n % 3
            if len(ring) <= slot or ring[slot][0].shape != xyz.shape or ring[slot][1].shape != features.shape:
                bufs = (torch.empty(xyz.shape, dtype=xyz.dtype).pin_memory(), torch.empty(features.shape, dtype=features.dtype).pin_memory(),
                        torch.empty((1,), dtype=torch.int32).pin_memory())
                if len(ring) <= slot:
                    ring.append(bufs)
                else:
                    ring[slot] = bufs
            h_xyz, h_feat, h_k = ring[slot]
            h_xyz.copy_(xyz.detach(), non_blocking=True)
            h_feat.copy_(features.detach(), non_blocking=True)
            if torch.is_tensor(k):
                h_k.copy_(k.reshape(1), non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            if pending is not None:
                finish(pending)
            pending = (f, h_xyz, h_feat, h_k if torch.is_tensor(k) else k, ev, slot)
        if pending is not None:
            finish(pending)
        for w in written:
            w.result()  # surface write errors
    return bin_files[lo:hi]
