"""Detect-and-describe pipeline: the north-star hot path as one call.

    FPS(num_clusters) -> gather_point -> query_ball_point (once) -> detector forward -> descriptor forward

All buffers are allocated once (180 GB of HBM: nothing is sized dynamically per call), every stage is one C-ABI
call on the caller's stream, and the whole sequence can be replayed from a CUDA graph (`use_graph=True`) so a
batch costs one launch from the host.  Mirrors Feat3dNet.get_inference_model (models/feat3dnet.py:258-313) in eval
mode; results are identical to calling the operators one by one.
"""
import importlib

import torch

_ROOT = __name__.split(".")[0]
_pfx = "3dfeatnet_b200." if _ROOT == "3dfeatnet_b200" else ""
_lib = importlib.import_module(_pfx + "_lib")
_f3d = importlib.import_module(_pfx + "models.feat3dnet")

STAGES = ("fps", "gather", "ball_query", "detector", "descriptor")


class DetectDescribePipeline:
    def __init__(self, batch, num_points, weights=None, num_clusters=512, radius=2.0, nsample=64, feature_dim=32,
                 no_regress=False, precision="fp32", device="cuda", use_graph=False, seed=0):
        self.B, self.N, self.M, self.S, self.F = batch, num_points, num_clusters, nsample, feature_dim
        self.radius, self.no_regress, self.precision = float(radius), no_regress, precision
        self.device = torch.device(device)
        self.L = _lib.lib()
        params = _f3d.init_params(seed, feature_dim, self.device) if weights is None else \
            _f3d.params_to_device(weights, self.device)
        self.packed = _f3d.fold_params(params, feature_dim)
        dev, B, N, M, S, F = self.device, batch, num_points, num_clusters, nsample, feature_dim
        self.xyz = torch.empty((B, N, 3), dtype=torch.float32, device=dev)
        self.fps_idx = torch.empty((B, M), dtype=torch.int32, device=dev)
        self.keypoints = torch.empty((B, M, 3), dtype=torch.float32, device=dev)
        self.idx = torch.empty((B, M, S), dtype=torch.int32, device=dev)
        self.pts_cnt = torch.empty((B, M), dtype=torch.int32, device=dev)
        self.attention = torch.empty((B, M), dtype=torch.float32, device=dev)
        self.orientation = torch.empty((B, M), dtype=torch.float32, device=dev)
        self.features = torch.empty((B, M, F), dtype=torch.float32, device=dev)
        self.fps_temp = torch.empty((B, N), dtype=torch.float32, device=dev) if N > 131072 else None
        self.bq_ws_bytes = self.L.f3d_query_ball_point_workspace_bytes(B, N)
        self.bq_ws = torch.empty((self.bq_ws_bytes,), dtype=torch.uint8, device=dev)
        self.ws_bytes = self.L.f3d_forward_workspace_bytes(B, M, F)
        self.ws = torch.empty((self.ws_bytes,), dtype=torch.uint8, device=dev)
        # pinned host mirrors for the end-to-end (host buffers in, host buffers out) entry point
        self.h_xyz = torch.empty((B, N, 3), dtype=torch.float32).pin_memory()
        self.h_out = torch.empty((B, M, 3 + 1 + 1 + F), dtype=torch.float32).pin_memory()
        self.d_out = torch.empty((B, M, 3 + 1 + 1 + F), dtype=torch.float32, device=dev)
        self.h2d_bytes = self.h_xyz.numel() * 4
        self.d2h_bytes = self.h_out.numel() * 4
        self.launches_per_step = None
        self._kp0, self._fps0 = self.keypoints, self.fps_idx  # the serial step's own buffers (step_pipelined re-points the public names)
        self.host_pipelined = False  # run_host_steps(): use the software-pipelined step
        self.host_ring = 4           # input buffers of the pipelined host loop (even, >= 4; bench.py sizes the ring beyond the L2)
        self._graph = None
        self._side = None  # second stream: ball-query grid build under FPS
        # ---- software-pipelined step (step_pipelined): sampling of batch i+1 beside the contractions of batch i ------------------
        # FPS is a chain of M dependent rounds per cloud: latency, one SM per cloud, the tensor pipe idle.  It therefore runs on
        # `fps_ctas` SMs (each CTA walking several clouds) while the persistent contraction kernels of the previous batch take the
        # other SMs (`det_sm_limit`) and the descriptor kernels, which start when the sampling has finished, all of them
        # (`desc_sm_limit`, 0 = no limit).  tune_pipelined() measures a few partitions and keeps the best.
        sms = torch.cuda.get_device_properties(self.device).multi_processor_count if self.device.type == "cuda" else 148
        self.num_sms = sms
        self.fps_ctas = (B + 1) // 2 if B >= 16 else B
        self.det_sm_limit = max(1, sms - self.fps_ctas) if self.fps_ctas < sms else 0
        self.desc_sm_limit = 0
        self._pl = None   # buffers / graphs of the pipelined step
        self._images_ready = False
        self._steady = False
        self.use_graph = use_graph
        # double buffers + side streams of the overlapped host<->device loop (run_host_steps)
        self._hp = None

    # one stage = one C-ABI call; `events` (optional list) receives a CUDA event after each stage
    def _enqueue(self, events=None, xyz=None):
        L, p, st = self.L, _lib.ptr, _lib.stream()
        xyz_buf = self.xyz if xyz is None else xyz
        B, N, M, S, F = self.B, self.N, self.M, self.S, self.F
        # after the first pass the workspace holds the tensor-core weight images of these (fixed) weights
        prec = _f3d.PRECISIONS[self.precision] | (_lib.PRECISION_IMAGES_CACHED if self._images_ready else 0)

        def mark():
            if events is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                events.append(e)

        mark()
        # The ball-query grid depends on the cloud only: outside the per-stage timing mode it is built on a side stream
        # while FPS (one CTA per cloud, 64 of the 148 SMs at the bench batch) selects the centres; inside a CUDA graph this
        # becomes a parallel branch.
        fork = events is None and N <= 262144
        if fork:
            if self._side is None:
                self._side = torch.cuda.Stream(device=self.device)
            cur = torch.cuda.current_stream()
            self._side.wait_stream(cur)
            with torch.cuda.stream(self._side):
                _lib.check(L.f3d_ball_grid_build(B, N, self.radius, p(xyz_buf), p(self.bq_ws), self.bq_ws_bytes, _lib.stream()),
                           "ball_grid_build")
        # sample_points (pointnet_common.py:14-29) = FPS + gather_point of the samples: one launch (the FPS kernel holds the
        # winner's coordinates each round); the "gather" stage of the per-stage timing is therefore empty
        _lib.check(L.f3d_farthest_point_sample_gather(B, N, M, p(xyz_buf), p(self.fps_temp), p(self.fps_idx), p(self.keypoints), st),
                   "fps+gather")
        mark()
        mark()
        if fork:
            cur.wait_stream(self._side)
            _lib.check(L.f3d_ball_grid_query(B, N, M, self.radius, S, p(xyz_buf), p(self.keypoints), p(self.idx),
                                             p(self.pts_cnt), p(self.bq_ws), self.bq_ws_bytes, st), "ball_grid_query")
        else:
            _lib.check(L.f3d_query_ball_point_ws(B, N, M, self.radius, S, p(xyz_buf), p(self.keypoints), p(self.idx),
                                                 p(self.pts_cnt), p(self.bq_ws), self.bq_ws_bytes, st), "query_ball_point")
        mark()
        _lib.check(L.f3d_detector_forward(B, N, M, S, self.radius, p(xyz_buf), p(self.keypoints), p(self.idx),
                                          p(self.packed), p(self.attention), p(self.orientation), prec, p(self.ws),
                                          self.ws_bytes, st), "detector_forward")
        mark()
        ori = None if self.no_regress else self.orientation
        _lib.check(L.f3d_descriptor_forward(B, N, M, S, self.radius, F, p(xyz_buf), p(self.keypoints), p(self.idx),
                                            p(ori), p(self.packed), p(self.features), prec, p(self.ws), self.ws_bytes, st),
                   "descriptor_forward")
        mark()

    # ---- software-pipelined step ------------------------------------------------------------------------------------------------
    def _enqueue_sample(self, xyz_buf, fps_idx, keypoints, bq_ws, max_ctas):
        """stage A of a batch: ball-query grid + farthest point sampling (+ gather of the samples)"""
        L, p, st = self.L, _lib.ptr, _lib.stream()
        _lib.check(L.f3d_ball_grid_build(self.B, self.N, self.radius, p(xyz_buf), p(bq_ws), self.bq_ws_bytes, st), "ball_grid_build")
        _lib.check(L.f3d_farthest_point_sample_gather_ctas(self.B, self.N, self.M, p(xyz_buf), p(self.fps_temp), p(fps_idx), p(keypoints),
                                                           int(max_ctas), st), "fps+gather")

    def _enqueue_contract(self, xyz_buf, keypoints, bq_ws, det_limit, desc_limit):
        """stage B of a batch: ball query against the prepared grid, detector, descriptor"""
        L, p, st = self.L, _lib.ptr, _lib.stream()
        B, N, M, S, F = self.B, self.N, self.M, self.S, self.F
        prec = _f3d.PRECISIONS[self.precision] | _lib.PRECISION_IMAGES_CACHED
        _lib.check(L.f3d_ball_grid_query(B, N, M, self.radius, S, p(xyz_buf), p(keypoints), p(self.idx), p(self.pts_cnt), p(bq_ws),
                                         self.bq_ws_bytes, st), "ball_grid_query")
        _lib.check(L.f3d_detector_forward(B, N, M, S, self.radius, p(xyz_buf), p(keypoints), p(self.idx), p(self.packed), p(self.attention),
                                          p(self.orientation), prec | _lib.precision_sm_limit(det_limit), p(self.ws), self.ws_bytes, st),
                   "detector_forward")
        ori = None if self.no_regress else self.orientation
        _lib.check(L.f3d_descriptor_forward(B, N, M, S, self.radius, F, p(xyz_buf), p(keypoints), p(self.idx), p(ori), p(self.packed),
                                            p(self.features), prec | _lib.precision_sm_limit(desc_limit), p(self.ws), self.ws_bytes, st),
                   "descriptor_forward")

    def _pipelined_state(self, ring=2):
        """ring: number of input buffers (2 when the batches are device resident, 4 when they arrive from the host)"""
        if self.N > 262144:
            raise _lib.F3DError("the pipelined step needs the ball-query grid (n <= 262144)")
        pl = self._pl
        if pl is None or pl["ring"] != ring:
            dev, B, N, M = self.device, self.B, self.N, self.M
            while not self._steady:
                self.step()  # builds the weight images and counts the launches of a serial step
            pl = dict(ring=ring, i=0, primed=False, graphs={}, side=torch.cuda.Stream(device=dev, priority=-1),
                      xyz=[self.xyz] + [torch.empty_like(self.xyz) for _ in range(ring - 1)],
                      kp=[self._kp0, torch.empty_like(self._kp0)], fps_idx=[self._fps0, torch.empty_like(self._fps0)],
                      bq_ws=[self.bq_ws, torch.empty_like(self.bq_ws)], pack=None)
            self._pl = pl
            if self._hp is not None:  # graphs of the host loop captured over the previous ring's buffers
                self._hp["graphs"] = {k: g for k, g in self._hp["graphs"].items() if not isinstance(k, tuple)}
        return pl

    def prime_pipelined(self, ring=2):
        """Prologue: sample the batch in input buffer 0 (its contractions run in the first pipelined step).  With ring=2 every
        input buffer is filled with the batch in self.xyz (device-resident benchmarking: the same batch every step)."""
        pl = self._pipelined_state(ring)
        if ring == 2:
            pl["xyz"][1].copy_(self.xyz)
        pl["i"] = 0
        self._enqueue_sample(pl["xyz"][0], pl["fps_idx"][0], pl["kp"][0], pl["bq_ws"][0], 0)
        pl["primed"] = True

    def _enqueue_pipelined(self, i, pack_to=None):
        """step i: contractions of batch i (input buffer i % ring) beside the sampling of batch i+1 (buffer (i+1) % ring)"""
        pl = self._pl
        r = pl["ring"]
        cur = torch.cuda.current_stream()
        pl["side"].wait_stream(cur)
        with torch.cuda.stream(pl["side"]):
            self._enqueue_sample(pl["xyz"][(i + 1) % r], pl["fps_idx"][(i + 1) & 1], pl["kp"][(i + 1) & 1], pl["bq_ws"][(i + 1) & 1], self.fps_ctas)
        self._enqueue_contract(pl["xyz"][i % r], pl["kp"][i & 1], pl["bq_ws"][i & 1], self.det_sm_limit, self.desc_sm_limit)
        if pack_to is not None:
            self._pack(pack_to, pl["kp"][i & 1])
        cur.wait_stream(pl["side"])

    def step_pipelined(self):
        """One pipelined step: afterwards attention / orientation / features / idx / pts_cnt (and self.keypoints / self.fps_idx, which
        are re-pointed) hold the results of the batch whose sampling ran in the previous step -- bit-identical to step() on that
        batch -- and the next batch has been sampled."""
        pl = self._pl
        if pl is None or not pl["primed"]:
            raise _lib.F3DError("call prime_pipelined() first")
        i = pl["i"]
        key = i % pl["ring"]
        if self.use_graph:
            g = pl["graphs"].get(key)
            if g is None:
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                # capture replays nothing: the sampled state of buffer (i+1) is produced by the first replay
                with torch.cuda.graph(g):
                    self._enqueue_pipelined(i)
                pl["graphs"][key] = g
            g.replay()
        else:
            self._enqueue_pipelined(i)
        self.keypoints, self.fps_idx = pl["kp"][i & 1], pl["fps_idx"][i & 1]
        pl["i"] = i + 1

    def set_partition(self, fps_ctas, det_sm_limit, desc_sm_limit):
        """SM partition of the pipelined step; drops the captured graphs."""
        self.fps_ctas, self.det_sm_limit, self.desc_sm_limit = int(fps_ctas), int(det_sm_limit), int(desc_sm_limit)
        if self._pl is not None:
            self._pl["graphs"] = {}
        if self._hp is not None:
            self._hp["graphs"] = {}

    def tune_pipelined(self, steps=8, candidates=None):
        """Measure a few SM partitions on the resident batch (CUDA events, `steps` pipelined steps each) and keep the fastest.
        Returns [(fps_ctas, det_sm_limit, desc_sm_limit, ms_per_step)]."""
        sms, B = self.num_sms, self.B
        if candidates is None:
            half, third = (B + 1) // 2, (B + 2) // 3
            candidates = [(B, 0, 0), (half, sms - half, 0), (half, sms - half, sms - half), (third, sms - third, 0),
                          (third, sms - third, sms - third), (half, 0, 0)]
            candidates = [c for c in dict.fromkeys(candidates) if c[0] >= 1 and c[1] >= 0 and c[0] <= sms]
        results = []
        for c in candidates:
            self.set_partition(*c)
            self.prime_pipelined()
            for _ in range(3):
                self.step_pipelined()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            s.record()
            for _ in range(steps):
                self.step_pipelined()
            e.record()
            torch.cuda.synchronize()
            results.append((c[0], c[1], c[2], s.elapsed_time(e) / steps))
        best = min(results, key=lambda r: r[3])
        self.set_partition(*best[:3])
        self.prime_pipelined()
        return results

    def step(self, events=None):
        """One pass over the batch already resident in self.xyz (device).  Results stay on the device."""
        self.keypoints, self.fps_idx = self._kp0, self._fps0
        if not self._steady:
            # pass 1 also builds the weight images (4 extra launches, once per set of weights); pass 2 is a steady-state
            # pass, whose launch count is what launches_per_step reports from then on
            self.L.f3d_reset_launch_count()
            self._enqueue()
            self.launches_per_step = int(self.L.f3d_launch_count())
            self._steady = self._images_ready
            self._images_ready = True
            return
        if self.use_graph and events is None:
            if self._graph is None:
                torch.cuda.synchronize()
                try:
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        self._enqueue()
                    self._graph = g
                except Exception as exc:  # capture is an optimisation only: run eagerly, loudly
                    import warnings

                    warnings.warn("CUDA graph capture failed (%s); running the stages eagerly" % exc)
                    self.use_graph = False
                    torch.cuda.synchronize()
                    self._enqueue()
                    return
            self._graph.replay()
        else:
            self._enqueue(events)

    def run(self, xyz):
        """xyz: (B,N,3) CUDA float32 -> dict of device tensors (views of the pipeline's buffers)."""
        _lib.require_cuda(xyz)
        self.xyz.copy_(xyz[:, :, :3])
        self.step()
        return dict(xyz=self.keypoints, features=self.features, attention=self.attention,
                    orientation=self.orientation, idx=self.idx, pts_cnt=self.pts_cnt, fps_idx=self.fps_idx)

    def step_host(self):
        """End to end: pinned host xyz (self.h_xyz) -> device -> pipeline -> packed [xyz|att|ori|desc] rows back in
        pinned host memory (self.h_out).  Asynchronous on the current stream; the caller synchronises."""
        self.xyz.copy_(self.h_xyz, non_blocking=True)
        self.step()
        self._pack(self.d_out)
        self.h_out.copy_(self.d_out, non_blocking=True)
        return self.h_out

    # ---- overlapped end-to-end loop: H2D of step i+1 and D2H of step i-1 run under the compute of step i ------------
    def _pack(self, d_out, keypoints=None):
        p = _lib.ptr
        kp = self.keypoints if keypoints is None else keypoints
        _lib.check(self.L.f3d_pack_rows(self.B * self.M, self.F, p(kp), p(self.attention), p(self.orientation),
                                        p(self.features), p(d_out), _lib.stream()), "pack_rows")

    def _host_pipe(self):
        if self._hp is None:
            dev, B, N, M, F = self.device, self.B, self.N, self.M, self.F
            hp = dict(
                s_h2d=torch.cuda.Stream(device=dev), s_comp=torch.cuda.Stream(device=dev), s_d2h=torch.cuda.Stream(device=dev),
                xyz=[torch.empty((B, N, 3), dtype=torch.float32, device=dev) for _ in range(2)],
                d_out=[torch.empty((B, M, 5 + F), dtype=torch.float32, device=dev) for _ in range(2)],
                h_out=[torch.empty((B, M, 5 + F), dtype=torch.float32).pin_memory() for _ in range(2)],
                graphs={0: None, 1: None})
            self._hp = hp
        return self._hp

    def run_host_steps(self, steps, host_batches=None, flush=None):
        """`steps` end-to-end passes through the public path with HOST buffers: every step copies its input batch from
        pinned host memory (host_batches[i % len], default self.h_xyz), computes, and copies the packed
        [xyz | attention | orientation | descriptor] rows back to pinned host memory -- all inside the timed region.
        Copies and compute of neighbouring steps overlap on three streams (double-buffered input / output).
        Returns (elapsed_ms by CUDA events from the first H2D to the last D2H, last host output)."""
        if self.host_pipelined:
            return self._run_host_steps_pipelined(steps, host_batches, flush)
        hp = self._host_pipe()
        host_batches = host_batches or [self.h_xyz]
        cur = torch.cuda.current_stream()
        for s in (hp["s_h2d"], hp["s_comp"], hp["s_d2h"]):
            s.wait_stream(cur)
        ev_h2d = [None, None]
        ev_comp = [None, None]
        ev_d2h = [None, None]
        start = torch.cuda.Event(enable_timing=True)
        end = torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(hp["s_h2d"]):
            start.record()
        for i in range(steps):
            b = i & 1
            with torch.cuda.stream(hp["s_h2d"]):
                if ev_comp[b] is not None:
                    hp["s_h2d"].wait_event(ev_comp[b])  # compute of step i-2 has consumed this input buffer
                hp["xyz"][b].copy_(host_batches[i % len(host_batches)], non_blocking=True)
                ev_h2d[b] = torch.cuda.Event()
                ev_h2d[b].record()
            with torch.cuda.stream(hp["s_comp"]):
                hp["s_comp"].wait_event(ev_h2d[b])
                if ev_d2h[b] is not None:
                    hp["s_comp"].wait_event(ev_d2h[b])  # D2H of step i-2 has drained this output buffer
                if flush is not None:
                    flush()
                if self.use_graph:
                    if hp["graphs"].get(b) is None:
                        raise _lib.F3DError("call warm_host_graphs() before timing with use_graph=True")
                    hp["graphs"][b].replay()
                else:
                    self._enqueue(xyz=hp["xyz"][b])
                    self._pack(hp["d_out"][b])
                ev_comp[b] = torch.cuda.Event()
                ev_comp[b].record()
            with torch.cuda.stream(hp["s_d2h"]):
                hp["s_d2h"].wait_event(ev_comp[b])
                hp["h_out"][b].copy_(hp["d_out"][b], non_blocking=True)
                ev_d2h[b] = torch.cuda.Event()
                ev_d2h[b].record()
        with torch.cuda.stream(hp["s_d2h"]):
            end.record()
        cur.wait_stream(hp["s_d2h"])
        cur.wait_stream(hp["s_comp"])
        cur.wait_stream(hp["s_h2d"])
        torch.cuda.synchronize()
        return start.elapsed_time(end), hp["h_out"][(steps - 1) & 1]

    def _run_host_steps_pipelined(self, steps, host_batches=None, flush=None):
        """run_host_steps with the software-pipelined step: step i runs [sampling of batch i+1] beside [contractions of batch i], so
        the H2D copy of batch i+1 has to land before step i starts (a ring of 4 input buffers keeps the copy of batch i+2 clear of the
        two buffers step i reads).  The timed region starts at the first H2D and includes the prologue (sampling of batch 0) and the
        sampling of one batch past the end; every batch is copied in from pinned host memory and its rows are copied back."""
        hp = self._host_pipe()
        R = self.host_ring
        if R < 4 or R & 1:
            raise _lib.F3DError("host_ring must be even and >= 4 (keypoints / grids are double-buffered by step parity)")
        pl = self._pipelined_state(ring=R)
        host_batches = host_batches or [self.h_xyz]
        cur = torch.cuda.current_stream()
        for st in (hp["s_h2d"], hp["s_comp"], hp["s_d2h"]):
            st.wait_stream(cur)
        ev_h2d, ev_comp, ev_d2h = {}, {}, {}
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

        def h2d(j):  # batch j -> input buffer j % R (last read by step j-R as batch and step j-R-1 as the next batch)
            with torch.cuda.stream(hp["s_h2d"]):
                if j - R in ev_comp:
                    hp["s_h2d"].wait_event(ev_comp[j - R])
                pl["xyz"][j % R].copy_(host_batches[j % len(host_batches)], non_blocking=True)
                ev_h2d[j] = torch.cuda.Event()
                ev_h2d[j].record()

        with torch.cuda.stream(hp["s_h2d"]):
            start.record()
        h2d(0)
        with torch.cuda.stream(hp["s_comp"]):  # prologue: sampling of batch 0
            hp["s_comp"].wait_event(ev_h2d[0])
            self._enqueue_sample(pl["xyz"][0], pl["fps_idx"][0], pl["kp"][0], pl["bq_ws"][0], 0)
        for i in range(steps):
            b = i & 1
            h2d(i + 1)  # (the batch after the last one is sampled and never used: the loop's fixed shape)
            with torch.cuda.stream(hp["s_comp"]):
                hp["s_comp"].wait_event(ev_h2d[i + 1])
                if i - 2 in ev_d2h:
                    hp["s_comp"].wait_event(ev_d2h[i - 2])  # D2H of step i-2 has drained this output buffer
                if flush is not None:
                    flush()
                if self.use_graph:
                    g = hp["graphs"].get(("pl", i % R))
                    if g is None:
                        raise _lib.F3DError("call warm_host_graphs() before timing with use_graph=True")
                    g.replay()
                else:
                    self._enqueue_pipelined(i, pack_to=hp["d_out"][b])
                ev_comp[i] = torch.cuda.Event()
                ev_comp[i].record()
            with torch.cuda.stream(hp["s_d2h"]):
                hp["s_d2h"].wait_event(ev_comp[i])
                hp["h_out"][b].copy_(hp["d_out"][b], non_blocking=True)
                ev_d2h[i] = torch.cuda.Event()
                ev_d2h[i].record()
        with torch.cuda.stream(hp["s_d2h"]):
            end.record()
        for st in (hp["s_d2h"], hp["s_comp"], hp["s_h2d"]):
            cur.wait_stream(st)
        torch.cuda.synchronize()
        return start.elapsed_time(end), hp["h_out"][(steps - 1) & 1]

    def warm_host_graphs(self):
        """Capture one CUDA graph per input/output buffer parity for run_host_steps (eager first, to set attributes)."""
        hp = self._host_pipe()
        if self.host_pipelined:
            R = self.host_ring
            pl = self._pipelined_state(ring=R)
            for j in range(R):
                pl["xyz"][j].copy_(self.h_xyz)
            with torch.cuda.stream(hp["s_comp"]):
                self._enqueue_sample(pl["xyz"][0], pl["fps_idx"][0], pl["kp"][0], pl["bq_ws"][0], 0)
                for i in range(R):
                    self._enqueue_pipelined(i, pack_to=hp["d_out"][i & 1])
            torch.cuda.synchronize()
            if self.use_graph:
                for i in range(R):
                    if hp["graphs"].get(("pl", i)) is None:
                        g = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(g, stream=hp["s_comp"]):
                            self._enqueue_pipelined(i, pack_to=hp["d_out"][i & 1])
                        hp["graphs"][("pl", i)] = g
            return
        while not self._steady:
            self.step()
        for b in range(2):
            hp["xyz"][b].copy_(self.h_xyz)
            with torch.cuda.stream(hp["s_comp"]):
                self._enqueue(xyz=hp["xyz"][b])
                self._pack(hp["d_out"][b])
            torch.cuda.synchronize()
            if self.use_graph and hp["graphs"].get(b) is None:
                try:
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=hp["s_comp"]):
                        self._enqueue(xyz=hp["xyz"][b])
                        self._pack(hp["d_out"][b])
                    hp["graphs"][b] = g
                except Exception as exc:
                    import warnings

                    warnings.warn("CUDA graph capture failed (%s); running the stages eagerly" % exc)
                    self.use_graph = False
                    torch.cuda.synchronize()
