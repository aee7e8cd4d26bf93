"""Drop-in for ramp.Ramp_vo.Ramp_vo (ramp/Ramp_vo.py:27-410): the online VO state machine.

Same constructor, `__call__(tstamp, input_tensor, intrinsics)`, `update()`, `keyframe()`,
`terminate()` and public buffers (`poses_`, `patches_`, `points_`, `colors_`, `n`, `m`, `ii`, `jj`,
`kk`) as the reference.  What changed underneath (HBM layout, DESIGN.md section "data layout"):

  * feature ring buffers are channels-last: fmap1_/fmap2_ are stored [mem,H,W,128], gmap_ is stored
    [mem*M,P,P,128]; the attributes keep the reference's logical shapes as permuted VIEWS, which is
    what routes `corr` to the tensor-core kernel (one launch for both pyramid levels, blend and
    permute fused) instead of 2 launches + ~12 elementwise kernels per level;
  * `reproject` is one fused kernel writing [1,E,2,P,P] directly (pops.transform + permute);
  * the graph bookkeeping of one update (neighbours, SoftAgg groups, BA patch numbering) is computed
    on the device once per graph change and cached — no `.to(kCPU)` round trip per update;
  * `point_cloud` after BA is one kernel over the patch centres.
Python keeps what the reference keeps in Python: ring-buffer index math, edge-list append/remove,
the motion model on two SE3 elements, keyframe decisions.
"""
import ctypes

import numpy as np
import torch
import torch.nn.functional as F

from . import _lib, altcorr, fastba, lietorch
from . import projective_ops as pops
from .lietorch import SE3
from .net import GraphPlans, VONet
from .vo_utils import copy_segments, filter_features, flatmeshgrid, preprocess_input, pyramid_level2


class _PatchifyGraph:
    """CUDA graph of `network.patchify` for the steady case (one event stack + one image per call):
    encoder -> /4 -> patch selection -> the four patch gathers -> pyramid level 2, ~450 kernel
    launches replayed as one.  Static input / output / recurrent-state buffers; the caller copies the
    outputs into the ring-buffer slot of the frame."""

    def __init__(self, vo):
        self.vo = vo
        dev, M, P = vo.device, vo.M, vo.P
        h4, w4 = vo.ht // vo.RES, vo.wd // vo.RES
        nb = vo.train_cfg["num_event_bins"]
        self.ev = torch.zeros(1, 1, nb, vo.ht, vo.wd, device=dev)
        self.im = torch.zeros(1, 1, 3, vo.ht, vo.wd, device=dev)
        self.mask = torch.tensor([True])
        self.gmap = torch.zeros(M, P, P, 128, device=dev, dtype=vo.fdtype)      # channels-last staging
        self.f1 = torch.zeros(h4, w4, 128, device=dev, dtype=vo.fdtype)
        self.f2 = torch.zeros(h4 // 4, w4 // 4, 128, device=dev, dtype=vo.fdtype)
        self.imap = torch.zeros(M, vo.DIM, device=dev, dtype=vo.fdtype)
        self.patches = torch.zeros(1, M, 3, P, P, device=dev)
        self.clr = torch.zeros(1, M, 3, device=dev)
        enc = vo.network.patchify.encoder
        self.state = None
        self.graph = None
        self.copy_stream = torch.cuda.Stream(device=dev)
        # the encoder graph replays on its OWN stream: frame t+1 is encoded while the recurrent update of frame t
        # (main stream) is still running — the two only meet at the staging buffers (events `replayed` / `consumed`)
        self.enc_stream = torch.cuda.Stream(device=dev)
        self.replayed = torch.cuda.Event()
        self.replayed.record(torch.cuda.current_stream(dev))
        self.consumed = torch.cuda.Event()
        self.consumed.record(torch.cuda.current_stream(dev))
        # parallel branches of the captured graph: context CNN and patch selection on side streams
        enc.branch_stream = torch.cuda.Stream(device=dev)
        vo.network.patchify.branch_stream = torch.cuda.Stream(device=dev)
        # spatial split of the GPU between the two streams (cfg.SM_SPLIT): the grids captured here fill `enc` SMs
        _lib.check(_lib.lib().rvo_set_sm_budget(vo.sm_split[0]), "rvo_set_sm_budget")
        # warm up on a side stream (cuDNN / cuBLAS handles, lazy module init), then capture
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        saved = list(enc.super_states)
        torch.backends.cudnn.benchmark = True    # fixed shapes: let cuDNN pick its fastest conv kernels
        with torch.cuda.stream(s):
            for _ in range(3):
                enc.super_states = [None, None, None]
                self._body(first=True)
            self.state = [t.detach().clone() for t in enc.super_states]
        torch.cuda.current_stream(dev).wait_stream(s)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        enc.super_states = list(self.state)
        l0 = _lib.lib().rvo_launch_count()
        with torch.cuda.graph(self.graph):
            self._body(first=False)
            for buf, new in zip(self.state, enc.super_states):
                buf.copy_(new)
        self.n_kernels = _lib.lib().rvo_launch_count() - l0   # librampvo kernels inside the graph
        self._held = _lib.Workspace.snapshot(dev)             # scratch the capture baked pointers to
        _lib.lib().rvo_set_sm_budget(0)
        enc.super_states = saved
        for b in self.state:
            b.zero_()

    @torch.no_grad()
    def _body(self, first):
        vo = self.vo
        gslot = self.gmap.permute(0, 3, 1, 2)[None]
        with torch.autocast("cuda", enabled=vo.autocast):
            fmap, gmap, imap, patches, _, clr = vo.network.patchify(
                input_=(self.ev, self.im, self.mask), patches_per_image=vo.M, event_bias=vo.event_bias,
                reinit_hidden=False, gmap_out=gslot)
        f = fmap[0, 0].permute(1, 2, 0)             # channels-last storage: this view is contiguous
        if not f.is_contiguous():
            f = f.contiguous()
        self.f1.copy_(f)
        self.f2.copy_(pyramid_level2(f if f.dtype == self.f2.dtype else f.to(self.f2.dtype)))
        self.imap.copy_(imap.view(vo.M, vo.DIM))
        self.patches.copy_(patches)
        self.clr.copy_(clr)

    def run(self, events, images, reinit):
        """launch the encoder graph of a new frame on the encoder stream; the caller waits for `replayed` on its own
        stream before it reads the staging buffers and records `consumed` once it has copied them out"""
        enc = self.vo.network.patchify.encoder
        dev = self.vo.device
        cur = torch.cuda.current_stream(dev)
        es = self.enc_stream
        es.wait_event(self.consumed)              # the previous frame's outputs have left the staging buffers
        with torch.cuda.stream(es):
            if reinit:
                for b in self.state:
                    b.zero_()
            else:   # pick up a state left by an eager (events-only) call
                for b, st in zip(self.state, enc.super_states):
                    if st is not None and st is not b:
                        es.wait_stream(cur)
                        b.copy_(st)
        if events.is_cuda and images.is_cuda:
            # device inputs are ordered after whatever the caller queued on its stream, unless it declares them
            # complete (vo.inputs_complete: e.g. frames pre-loaded and synchronised) — only then can this frame's
            # encoder overlap the previous frame's update
            if not self.vo.inputs_complete:
                es.wait_stream(cur)
            with torch.cuda.stream(es):
                self.ev.copy_(events)
                self.im.copy_(images)
        else:
            # host inputs (pinned memory copies asynchronously): H->D on a copy stream
            cs = self.copy_stream
            cs.wait_event(self.replayed)          # the static input buffers are free once the last replay is done
            with torch.cuda.stream(cs):
                self.ev.copy_(events, non_blocking=True)
                self.im.copy_(images, non_blocking=True)
            es.wait_stream(cs)
        with torch.cuda.stream(es):
            self.graph.replay()
            self.replayed.record(es)
        self.vo.graph_kernel_launches += self.n_kernels
        enc.super_states = list(self.state)


class _UpdateGraph:
    """CUDA graph of one recurrent update for a fixed edge count E and window length: static copies
    of ii/jj/kk, the graph-plan sorts, reproject, corr, the update operator and both BA iterations.
    The sliding window start t0 is a device scalar (rvo_ba_forward_dyn)."""

    def __init__(self, vo, E, n_free):
        self.vo = vo
        dev = vo.device
        self.ii = torch.zeros(E, dtype=torch.long, device=dev)
        self.jj = torch.zeros(E, dtype=torch.long, device=dev)
        self.kk = torch.zeros(E, dtype=torch.long, device=dev)
        self.t0 = torch.zeros(1, dtype=torch.int32, device=dev)
        self.n_free = n_free
        vo._net_join()                           # the warm-up / capture below read vo.net on this stream
        self.side = torch.cuda.Stream(device=dev)
        self.mid_event = torch.cuda.Event(external=True)
        self.net_in = vo.net                     # view of the current ping-pong buffer
        self.net_out = vo._net_other(E)
        self._load(vo.n - n_free)
        vo.corr_tiles(vo.reproject())            # sizes the persistent corr buffer outside the capture
        snap = (vo.poses_.clone(), vo.patches_.clone(), self.net_in.clone())
        _lib.check(_lib.lib().rvo_set_sm_budget(vo.sm_split[1]), "rvo_set_sm_budget")
        try:
            # the capture stream is where the replays' scratch lives: warm up ON it so that every workspace
            # (keyed by stream) has its final size before pointers are baked in
            self.graph = torch.cuda.CUDAGraph()
            if vo._capture_stream is None:       # one per Ramp_vo: its graphs share scratch, instances do not
                vo._capture_stream = torch.cuda.Stream(device=dev)
            s = vo._capture_stream
            s.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(s):
                for _ in range(2):               # warm-up (workspaces, cuBLAS handles) on a side stream
                    self._body()
            torch.cuda.current_stream(dev).wait_stream(s)
            torch.cuda.synchronize(dev)
            vo.poses_.copy_(snap[0]); vo.patches_.copy_(snap[1]); self.net_in.copy_(snap[2])
            l0 = _lib.lib().rvo_launch_count()
            with torch.cuda.graph(self.graph, stream=s):
                self._body()
            self.n_kernels = _lib.lib().rvo_launch_count() - l0   # librampvo kernels inside the graph
        finally:
            _lib.lib().rvo_set_sm_budget(0)
            # the warm-up / capture passes ran BA on the live state: always put it back
            vo.poses_.copy_(snap[0]); vo.patches_.copy_(snap[1]); self.net_in.copy_(snap[2])
        # every buffer the capture baked a raw pointer to and does not own stays alive with the graph: the
        # correlation rows, both hidden-state buffers and the library scratch (they are replaced, not resized
        # in place, when they grow — a cached graph then keeps using its own, still valid, copies)
        self._held = (vo._corrt_buf, vo._corr_buf, tuple(vo._net_bufs), _lib.Workspace.snapshot(dev))

    def _load(self, t0):
        vo = self.vo
        copy_segments([(vo.ii, self.ii), (vo.jj, self.jj), (vo.kk, self.kk)])      # one launch
        self.t0.fill_(t0)

    def _body(self):
        vo = self.vo
        # the graph-plan sorts depend only on the edge list: a parallel branch next to reproject + corr
        cur = torch.cuda.current_stream(vo.device)
        side = self.side
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            plans = vo._new_plans(self.ii, self.jj, self.kk)
        self.plans = plans
        def after_corr():
            cur.wait_stream(side)
            # the hidden-state rows of the new edge list (rvo_net_rows on vo._net_stream) are only needed from here
            # on: an external event-wait node lets reproject + corr start before that gather has finished
            cur.wait_event(vo._net_ready)
            # an EXTERNAL event node: the encoder stream of the next frame can wait for "reproject + corr of this
            # update are done" (Ramp_vo.encoder_after_corr) — the latency-critical head of the update then has the
            # GPU to itself
            self.mid_event.record(cur)
        _, self.weight = vo._update_body(self.ii, self.jj, self.kk, self.net_in, self.net_out, plans,
                                         0, self.n_free, t0_dev=self.t0, before_update=after_corr,
                                         with_ba=vo.world_size == 1)

    def run(self, t0):
        vo = self.vo
        if not vo._net_ready_fresh:          # hidden state produced on this stream (no side-stream gather pending)
            vo._net_ready.record(torch.cuda.current_stream(vo.device))
        vo._net_ready_fresh = False
        self._load(t0)
        self.graph.replay()
        self.vo.graph_kernel_launches += self.n_kernels


class Ramp_vo:
    def __init__(self, cfg, network, train_cfg, ht=480, wd=640, device="cuda", use_graphs=True, pipeline=False,
                 world_size=1, rank=0, group=None):
        """world_size > 1 (extension): the patch graph is SHARDED across the ranks of a
        torch.distributed group by source frame — every rank sees the same frames and keeps the full pose /
        patch / feature state, but only the edges of the frames it owns: reproject, corr, the update operator and
        the patch blocks of BA are rank-local; per Gauss-Newton iteration the reduced camera system [S | y] is
        all-reduced (NCCL over NVLink), every rank solves it identically, and the refined depths of the window
        are exchanged once per update."""
        self.cfg = cfg
        self.world_size, self.rank, self.group = int(world_size), int(rank), group
        self._owner = []            # owner rank of every live frame slot (follows frames through keyframe drops)
        self._owner_dev = None      # the same table on the device (sharded mode)
        self.collective_bytes = 0   # bytes this rank put through collectives (sharded mode)
        self.collective_calls = 0
        self.event_bias = train_cfg["event_bias"]
        self.train_cfg = train_cfg
        self.device = torch.device(device)
        self.lmbda = torch.as_tensor([1e-4], device=self.device)
        self.load_weights(network)
        self.is_initialized = False
        self.enable_timing = False

        self.n = 0      # number of frames
        self.m = 0      # number of patches
        self.M = self.cfg.PATCHES_PER_FRAME
        self.N = self.cfg.BUFFER_SIZE
        self.ht, self.wd = ht, wd
        DIM, RES, P, M, N = self.DIM, self.RES, self.P, self.M, self.N
        dev = self.device

        self.tlist = []
        # (encoder SMs, update SMs) for the two CUDA graphs of the pipelined frame; 0 = the whole GPU
        # Scheduling of the pipelined frame (measured on B200, default.yaml; profiles/r02_frame_scheduling.md):
        #   SM_SPLIT            the encoder graph of frame t+1 and the update graph of frame t are captured with
        #                       disjoint SM budgets (rvo_set_sm_budget): they run side by side instead of time-slicing
        #                       the GPU kernel by kernel;
        #   ENCODER_AFTER_CORR  the encoder of frame t+1 starts once reproject + corr of update t are done;
        #   ENCODER_IN_GAP      (alternative) the encoder of frame t+1 starts when update t has finished.
        split = getattr(cfg, "SM_SPLIT", None)
        self.sm_split = tuple(split) if split is not None else (
            self._auto_sm_split(cfg, tiles=-(-(ht // 4) * (wd // 4) // 128)) if pipeline and world_size == 1 else (0, 0))
        self.encoder_in_gap = bool(getattr(cfg, "ENCODER_IN_GAP", False))
        self.encoder_after_corr = bool(getattr(cfg, "ENCODER_AFTER_CORR", True))
        self._last_ugraph = None
        self.fast_edges = bool(getattr(cfg, "FAST_EDGES", True))   # fused patch-graph step (rvo_edges_step)
        self._pending_lim = None
        self._pending_drop = -1     # keyframe whose removal rides on the next rvo_edges_step
        self._net_stream = None
        self._net_ready = torch.cuda.Event(external=True) if self.device.type == "cuda" else None
        self._net_ready_fresh = False
        self._edge_status = None
        self._edge_src = None
        self._edge_tiles = None
        self._edge_epoch = 0
        self._min_src = 0
        self._frame_done = None
        self.last_weight = None
        self.patch_dict_ = None          # pose-prediction caches (Ramp_vo.py:94-95)
        self.patches_models = None
        self.counter = 0
        self.tstamps_ = torch.zeros(N, dtype=torch.long, device=dev)
        self.poses_ = torch.zeros(N, 7, dtype=torch.float, device=dev)
        self.patches_ = torch.zeros(N, M, 3, P, P, dtype=torch.float, device=dev)
        self.intrinsics_ = torch.zeros(N, 4, dtype=torch.float, device=dev)
        self.points_ = torch.zeros(N * M, 3, dtype=torch.float, device=dev)
        self.colors_ = torch.zeros(N, M, 3, dtype=torch.uint8, device=dev)
        self.index_ = torch.zeros(N, M, dtype=torch.long, device=dev)
        self.index_map_ = torch.zeros(N, dtype=torch.long, device=dev)

        self.mem = 32
        self.autocast = bool(self.cfg.MIXED_PRECISION)
        self.fdtype = torch.half if self.autocast else torch.float
        self.kwargs = {"device": dev, "dtype": self.fdtype}
        h4, w4 = ht // RES, wd // RES
        self.imap_ = torch.zeros(self.mem, M, DIM, **self.kwargs)
        # channels-last storage, reference-shaped views
        self._gmap_store = torch.zeros(self.mem * M, P, P, 128, **self.kwargs)
        self._fmap1_store = torch.zeros(self.mem, h4, w4, 128, **self.kwargs)
        self._fmap2_store = torch.zeros(self.mem, h4 // 4, w4 // 4, 128, **self.kwargs)
        self.gmap_ = self._gmap_store.permute(0, 3, 1, 2).view(self.mem, M, 128, P, P)
        self.fmap1_ = self._fmap1_store.permute(0, 3, 1, 2)[None]
        self.fmap2_ = self._fmap2_store.permute(0, 3, 1, 2)[None]
        self.pyramid = (self.fmap1_, self.fmap2_)

        # hidden state of the edges: two ping-pong buffers [1, capacity, DIM]; self.net is a view
        self._net_bufs = [None, None]
        self._net_cur = 0
        self.net = torch.zeros(1, 0, DIM, device=dev, dtype=torch.float)
        self.ii = torch.as_tensor([], dtype=torch.long, device=dev)
        self.jj = torch.as_tensor([], dtype=torch.long, device=dev)
        self.kk = torch.as_tensor([], dtype=torch.long, device=dev)
        self._plans = None          # GraphPlans of the current edge list
        self._pair_cnt = None       # host-side {(source frame, target frame): number of edges}
        self.use_graphs = use_graphs
        # pipeline=True: the keyframe step of frame t (whose decision needs a device->host read) is finished at
        # the start of the next __call__, after that frame's encoder graph has been launched, so the GPU runs
        # encoder(t+1) while the host does the edge bookkeeping of frame t.  Same work, same results; between
        # calls the public state is the one BEFORE the keyframe step until sync() / terminate() / the next call.
        self.pipeline = pipeline
        # True: device-resident input tensors handed to __call__ are already complete (produced and synchronised
        # before the call), so the encoder stream need not wait for the work queued on the caller's stream
        self.inputs_complete = False
        self._pending_kf = None     # (pinned host buffer, event) of a keyframe step that was begun
        self._pgraph = None         # _PatchifyGraph, captured at the first frame
        self._corr_buf = None       # [1, capacity, 896] correlation rows (882 used)
        self._corrt_buf = None      # [1, capacity, 1008] correlation rows in the tile layout
        self._ugraphs = {}          # (E, window, buffer parity) -> _UpdateGraph
        self._capture_stream = None # warm-up + capture stream of the update graphs (keys their scratch buffers)
        self._ukey_prev, self._ukey_hist = None, []
        self.graph_kernel_launches = 0   # librampvo kernels executed through CUDA-graph replays

        self.poses_[:, 6] = 1.0
        self.delta = {}
        self.Id = SE3.Identity(1, device=dev)

    # ------------------------------------------------------------------ weights
    def load_weights(self, network):
        """ramp/Ramp_vo.py:103-129: a path to a reference checkpoint or a VONet instance."""
        if isinstance(network, str):
            ckpt = torch.load(network, map_location="cpu")
            sd = ckpt.get('model_state_dict') or ckpt
            sd = {k.replace('module.', ''): v for k, v in sd.items() if "update.lmbda" not in k}
            self.network = VONet(cfg=self.train_cfg)
            self.network.load_state_dict(sd)
        else:
            self.network = network
        self.DIM = self.network.DIM
        self.RES = self.network.RES
        self.P = self.network.P
        self.network.to(self.device)
        self.network.eval()

    # ------------------------------------------------------------------ views (Ramp_vo.py:131-157)
    @property
    def poses(self):
        return self.poses_.view(1, self.N, 7)

    @property
    def patches(self):
        return self.patches_.view(1, self.N * self.M, 3, self.P, self.P)

    @property
    def intrinsics(self):
        return self.intrinsics_.view(1, self.N, 4)

    @property
    def ix(self):
        return self.index_.view(-1)

    @property
    def imap(self):
        return self.imap_.view(1, self.mem * self.M, self.DIM)

    @property
    def gmap(self):
        return self._gmap_store.permute(0, 3, 1, 2)[None]      # [1, mem*M, 128, P, P] view

    # ------------------------------------------------------------------ trajectory (Ramp_vo.py:155-173)
    def get_pose(self, t):
        if t in self.traj:
            return SE3(self.traj[t])
        t0, dP = self.delta[t]
        return dP * self.get_pose(t0)

    def terminate(self):
        """interpolate missing poses; returns (poses [counter,7] camera-to-world, tstamps)."""
        self.sync()
        self.traj = {}
        ts = self.tstamps_[:self.n].tolist()
        for i in range(self.n):
            self.traj[ts[i]] = self.poses_[i]
        poses = [self.get_pose(t) for t in range(self.counter)]
        poses = lietorch.stack(poses, dim=0)
        poses = poses.inv().data.cpu().numpy()
        return poses, np.array(self.tlist, dtype=float)

    @staticmethod
    def _auto_sm_split(cfg, tiles=150, n_sms=148):
        """(encoder SMs, update SMs) for the pipelined frame: the split that balances the two streams in a two-line
        cost model fitted to the B200 measurements of profiles/r02_frame_scheduling.md — the encoder is latency-bound
        (time ~ rounds of its 150-tile layers: flat down to 74 SMs, then ~ 1 / SMs), the update is half tensor-bound,
        half HBM-bound and scales with the steady-state edge count of the preset.  default.yaml -> (30, 118),
        fast.yaml -> (50, 98), precise.yaml -> no split; cfg.SM_SPLIT overrides it."""
        edges = 0.86 * cfg.PATCHES_PER_FRAME * (2 * cfg.PATCH_LIFETIME - 1) * cfg.REMOVAL_WINDOW
        t_u = 1.14 * edges / 45312.0
        if t_u > 4.0:       # precise.yaml-sized graphs: the encoder is noise next to the update — measured 73.5 vs 71.2
            return 0, 0     # frames/s without / with a split
        best = None
        for se in range(8, 76, 4):
            t_enc = 0.58 * (0.25 + 0.75 * max(1.0, 74.0 / se))
            t_upd = t_u * (0.45 + 0.55 * n_sms / (n_sms - se)) + 0.25
            cost = max(t_enc, t_upd)
            if best is None or cost < best[0] - 1e-9:
                best = (cost, se)
        se = best[1] + 4       # a starved encoder stalls the whole frame, a starved update only loses its share
        # wave alignment: most encoder layers run at 1/4 resolution in tiles of 128 pixels, one tile per CTA per round
        # (150 tiles at 480x640): 30 SMs do them in 5 rounds, 28 SMs need 6 (measured 572 vs 564 frames/s)
        up = -(-tiles // max(tiles // se, 1))
        se = up if up <= se + 8 else -(-tiles // -(-tiles // se))
        return se, n_sms - se

    # ------------------------------------------------------------------ pose prediction (Ramp_vo.py:412-545)
    def _virtual_frame(self, last_keyframe_number):
        """the common head of both prediction entry points: clone the state, bootstrap the pose of the virtual frame
        `last_keyframe_number` (0-based index) with the motion model, connect every patch of the last
        PATCH_LIFETIME - 1 frames to it and reproject (Ramp_vo.py:415-443 == :449-475)"""
        from . import pose_prediction as PP
        self.sync()
        next_frame_number = last_keyframe_number + 1
        next_frame_index = next_frame_number - 1
        poses = self.poses.clone()
        poses[:, next_frame_index] = PP.motion_bootstrap(poses=poses[0], n=self.n, MOTION_MODEL=self.cfg.MOTION_MODEL,
                                                         MOTION_DAMPING=self.cfg.MOTION_DAMPING)
        intrinsics = self.intrinsics.clone()
        intrinsics[:, next_frame_index] = intrinsics[:, next_frame_index - 1]
        patches = self.patches.clone()
        if self.last_weight is None:
            raise RuntimeError("pose prediction needs at least one update() (last_weight is empty)")
        weights = self.last_weight.clone().float()
        ii, jj, kk, weights_up = PP.add_forward_elements(
            frame_num=next_frame_number, patch_extracted_num=self.M, ii=self.ii, jj=self.jj, kk=self.kk, ix=self.ix,
            r=self.cfg.PATCH_LIFETIME, weights=weights)
        coords = self.reproject(indicies=(ii, jj, kk), poses=poses, patches=patches, intrinsics=intrinsics)
        return next_frame_number, next_frame_index, poses, patches, intrinsics, ii, jj, kk, weights_up, coords

    @torch.no_grad()
    def efficient_pose_prediction(self, sec_to_pred_future, abs_time, last_keyframe_number, deg=3, frequency=30):
        """Ramp_vo.py:414-443: motion-model bootstrap of the virtual frame only (the reference computes the
        reprojection and discards it; the state is not modified)"""
        return self._virtual_frame(last_keyframe_number)[2][:, last_keyframe_number]

    @torch.no_grad()
    def predict_future_pose(self, sec_to_pred_future, abs_time, last_keyframe_number, deg=3, frequency=30):
        """Ramp_vo.py:446-514: extrapolate every patch track `sec_to_pred_future` frames ahead with per-patch
        smoothing splines (fitted once and cached, like the reference's patch_dict_ / patches_models), use the
        predictions as BA targets of a virtual frame behind the last keyframe, run 2 Gauss-Newton iterations on
        CLONES of the state and append the virtual frame's pose to the trajectory (update_attributes).
        Deviation: the reference hands the whole [1,E,2,3,3] coordinate tensor to cuda_ba, whose `view(-1, 2)`
        (ba_cuda.cu:462) then reads the first E/9 edges' grids as E targets; here the patch-centre pixel of every
        edge is the target, as in update() (Ramp_vo.py:288)."""
        from . import pose_prediction as PP
        (next_frame_number, next_frame_index, poses, patches, intrinsics, ii, jj, kk, weights_up,
         coords) = self._virtual_frame(last_keyframe_number)
        if self.patch_dict_ is None:
            self.patch_dict_ = PP.compute_patch_track(coords=coords, ii=ii, jj=jj, kk=kk, image_to_proj=next_frame_index)
        if self.patches_models is None:
            self.patches_models = PP.fit_model_patch_track(
                next_frame_index=next_frame_index, patch_dict=self.patch_dict_, img_to_keyframe_map=self.tstamps_,
                ii=ii, jj=jj, data_shape=(self.ht, self.wd), frequency=frequency, deg=deg)
        coords, updated_weight = PP.predict_patch_on_model(
            patch_models=self.patches_models, step_to_pred_future=sec_to_pred_future, frequency=frequency,
            next_frame_index=next_frame_index, coords=coords.contiguous(), weights=weights_up, ii=ii, jj=jj, kk=kk)
        target = coords[:, :, :, self.P // 2, self.P // 2].contiguous()
        t0 = max(next_frame_number - self.cfg.OPTIMIZATION_WINDOW if self.is_initialized else 1, 1)
        t1 = next_frame_number
        try:
            fastba.BA(poses, patches, intrinsics, target, updated_weight.contiguous(), self.lmbda, ii, jj, kk, t0, t1,
                      self.M, 2, eff_impl=False)
        except RuntimeError as e:
            print(f"WARNING: BA failed...{e}")
        self.update_attributes(abs_time=abs_time, next_frame_index=next_frame_index, poses=poses)

    def update_attributes(self, abs_time, next_frame_index, poses):
        """make the virtual frame visible to terminate() (Ramp_vo.py:517-524)"""
        assert int(self.tstamps_[self.n - 1].item()) != 0
        self.tstamps_[self.n] = abs_time
        self.poses_[self.n] = poses[0, next_frame_index]
        self.tlist.append(abs_time)
        self.counter += 1
        self.n += 1

    def remove_attributes(self):
        """undo update_attributes (Ramp_vo.py:526-533)"""
        self.n -= 1
        self.counter -= 1
        self.tlist.pop()
        self.poses_[self.n] = torch.zeros(7, dtype=torch.float, device=self.device)
        self.poses_[:, 6] = 1.0
        self.tstamps_[self.n] = 0

    # ------------------------------------------------------------------ hot-path pieces
    def corr(self, coords, indicies=None):
        """local correlation volume [1,E,882] (Ramp_vo.py:175-182), one fused launch"""
        ii, jj = indicies if indicies is not None else (self.kk, self.jj)
        E = ii.numel()
        # rows padded to 896 halves (pad columns stay zero) so the first GEMM of the update operator
        # has K % 8 == 0; the returned tensor is the [1,E,882] view of it
        buf = self._corr_buf
        if buf is None or buf.shape[1] < E:
            buf = self._corr_buf = torch.zeros(1, max(E, 1) * 5 // 4 + 256, 896, dtype=self.fdtype,
                                               device=self.device)
        out = buf[:, :E, :882]
        return altcorr.corr_pyramid(self.gmap, self.pyramid, coords, ii, jj, self.M * self.mem,
                                    self.mem, 3, out=out)

    def corr_tiles(self, coords, indicies=None):
        """the same correlation volume in the tile layout of altcorr.corr_tiles ([1,E,1008]), computed
        on the tcgen05 tensor cores; the update operator consumes it with permuted first-layer weights"""
        ii, jj = indicies if indicies is not None else (self.kk, self.jj)
        E = ii.numel()
        buf = self._corrt_buf
        if buf is None or buf.shape[1] < E:
            buf = self._corrt_buf = torch.zeros(1, max(E, 1) * 5 // 4 + 256, 18 * altcorr.TILE_GROUP,
                                                dtype=self.fdtype, device=self.device)
        return altcorr.corr_tiles(self.gmap, self.pyramid, coords, ii, jj, self.M * self.mem, self.mem,
                                  out=buf[:, :E])

    def reproject(self, indicies=None, poses=None, patches=None, intrinsics=None):
        """reproject patch k from i -> j: coords [1,E,2,P,P] (Ramp_vo.py:184-192)"""
        (ii, jj, kk) = indicies if indicies is not None else (self.ii, self.jj, self.kk)
        poses = poses if poses is not None else self.poses
        patches = patches if patches is not None else self.patches
        intrinsics = intrinsics if intrinsics is not None else self.intrinsics
        return pops.reproject_cf(SE3(poses), patches, intrinsics, ii, jj, kk)

    def _net_reserve(self, E_new):
        """make both hidden-state buffers hold at least E_new edges, keeping self.net's content"""
        cur = self._net_bufs[self._net_cur]
        if cur is not None and cur.shape[1] >= E_new:
            return
        cap = max(E_new * 5 // 4, 4096)
        E0 = self.net.shape[1]
        new = [torch.zeros(1, cap, self.DIM, device=self.device, dtype=torch.float) for _ in range(2)]
        new[0][:, :E0] = self.net
        self._net_bufs, self._net_cur = new, 0
        self.net = new[0][:, :E0]

    def _net_other(self, E):
        """the [1,E,DIM] view of the buffer self.net does NOT live in (output of the next stage)"""
        self._net_reserve(max(E, self.net.shape[1]))
        return self._net_bufs[1 - self._net_cur][:, :E]

    def _net_swap(self, E):
        self._net_cur = 1 - self._net_cur
        self.net = self._net_bufs[self._net_cur][:, :E]

    # The edge lists live on the device (self.ii / jj / kk, public like the reference's).  The host
    # keeps only the number of edges per (source frame, target frame) pair — enough to know how many
    # edges a removal keeps, so that the compaction can use torch.nonzero_static instead of a boolean
    # index that waits for the device.
    def _pair_counts(self):
        pc = self._pair_cnt
        if pc is None or sum(pc.values()) != self.ii.numel():      # edited from outside: recount once
            key = (self.ii * (self.N + 1) + self.jj).cpu().numpy()
            u, c = np.unique(key, return_counts=True)
            pc = self._pair_cnt = {(int(k) // (self.N + 1), int(k) % (self.N + 1)): int(v) for k, v in zip(u, c)}
            self._min_src = 0
        return pc

    def append_factors(self, ii, jj, pairs=None):
        """add factors to the graph (Ramp_vo.py:194-201); new edges start with a zero hidden state.
        pairs: {(i, j): count} of the new edges when the caller knows it (no device read needed)"""
        self._net_join()
        pc = self._pair_counts()
        if pairs is None:
            key = (torch.div(ii, self.M, rounding_mode="floor") * (self.N + 1) + jj).cpu().numpy()
            u, c = np.unique(key, return_counts=True)
            pairs = {(int(k) // (self.N + 1), int(k) % (self.N + 1)): int(v) for k, v in zip(u, c)}
        for k, v in pairs.items():
            pc[k] = pc.get(k, 0) + v
        self.jj = torch.cat([self.jj, jj])
        self.kk = torch.cat([self.kk, ii])
        self.ii = torch.cat([self.ii, torch.div(ii, self.M, rounding_mode="floor")])
        E0, n = self.net.shape[1], len(ii)
        self._net_reserve(E0 + n)
        buf = self._net_bufs[self._net_cur]
        buf[:, E0:E0 + n].zero_()
        self.net = buf[:, :E0 + n]
        self._plans = None

    def remove_factors(self, m, pair_pred=None):
        """remove factors from the graph (Ramp_vo.py:203-208).  m: boolean device mask (reference
        signature).  pair_pred(i, j) -> bool: the same predicate on (source, target) frame pairs; with
        it the number of surviving edges is known on the host and the compaction does not synchronise"""
        self._net_join()
        pc = self._pair_counts()
        if pair_pred is None:
            n_keep = int((~m).sum().item())
            self._pair_cnt = None                                   # recount lazily
        else:
            gone = [k for k in pc if pair_pred(*k)]
            for k in gone:
                del pc[k]
            n_keep = sum(pc.values())
        if n_keep == m.numel():
            return
        keep = torch.nonzero_static(~m, size=n_keep).view(-1)
        self.ii = self.ii[keep]
        self.jj = self.jj[keep]
        self.kk = self.kk[keep]
        out = self._net_other(keep.numel())
        torch.index_select(self.net, 1, keep, out=out)
        self._net_swap(keep.numel())
        self._plans = None

    def _flush_pending_drop(self):
        """apply a keyframe drop that was left to rvo_edges_step with plain tensor ops (Ramp_vo.py:249-262)"""
        k, self._pending_drop = self._pending_drop, -1
        if k < 0:
            return
        pc, self._pair_cnt = self._pair_cnt, None       # remove_factors wants counts that match the device lists
        m = (self.ii == k) | (self.jj == k)
        n_keep = sum(pc.values())
        self._net_join()
        if n_keep != m.numel():
            keep = torch.nonzero_static(~m, size=n_keep).view(-1)
            self.ii, self.jj, self.kk = self.ii[keep], self.jj[keep], self.kk[keep]
            out = self._net_other(n_keep)
            torch.index_select(self.net, 1, keep, out=out)
            self._net_swap(n_keep)
            self._plans = None
        gi = self.ii > k
        self.kk = self.kk - gi * self.M
        self.ii = self.ii - gi.to(self.ii.dtype)
        self.jj = self.jj - (self.jj > k).to(self.jj.dtype)
        self._pair_cnt = pc

    def _net_join(self):
        """make the current stream wait for a hidden-state gather still running on the side stream (_edges_step)"""
        if self._net_ready_fresh:
            torch.cuda.current_stream(self.device).wait_event(self._net_ready)
            self._net_ready_fresh = False

    def _edges_step(self, lim):
        """remove_factors(ii < lim) + append_factors(forward) + append_factors(backward) of the frame that was just
        added (self.n already counts it) in two launches: the edge lists and the hidden-state rows are rebuilt on the
        device by rvo_edges_step, the host only keeps its pair counts in step."""
        drop_k, self._pending_drop = self._pending_drop, -1
        pc = self._pair_cnt if drop_k >= 0 else self._pair_counts()    # (a pending drop already left the host counts)
        r, M, n = self.cfg.PATCH_LIFETIME, self.M, self.n
        E0 = self.ii.numel()
        n_removed = E0 - sum(pc.values())                              # edges of the dropped keyframe
        for i in range(self._min_src, max(lim, self._min_src)):      # source frames that leave the window
            for j in range(max(i - r - 1, 0), i + r + 2):
                n_removed += pc.pop((i, j), 0)
        self._min_src = max(lim, self._min_src)
        f0, f1, j0 = max(n - r, 0), max(n - 1, 0), max(n - r, 0)
        for i in range(f0, f1):
            pc[(i, n - 1)] = pc.get((i, n - 1), 0) + M
        for j in range(j0, n):
            pc[(n - 1, j)] = pc.get((n - 1, j), 0) + M
        E1 = E0 - n_removed + M * (f1 - f0) + M * (n - j0)
        dev = self.device
        if self._edge_status is None:
            self._edge_status = torch.zeros(1, dtype=torch.float32, device=dev)
        if self._edge_src is None or self._edge_src.numel() < E1:
            self._edge_src = torch.empty(max(E1 * 5 // 4, 4096), dtype=torch.int32, device=dev)
        nt = int(_lib.lib().rvo_edges_step_tiles(E0))
        if self._edge_tiles is None or self._edge_tiles.numel() < nt:
            self._edge_tiles = torch.zeros(max(2 * nt, 256), dtype=torch.int64, device=dev)
        self._edge_epoch = (self._edge_epoch % 0x7fffffff) + 1
        ii, jj, kk = (torch.empty(E1, dtype=torch.long, device=dev) for _ in range(3))
        out = self._net_other(E1)
        cur = torch.cuda.current_stream(dev)
        if self._net_stream is None:
            self._net_stream = torch.cuda.Stream(device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().rvo_edges_step(
                _lib.ptr(self.ii), _lib.ptr(self.jj), _lib.ptr(self.kk), E0, int(drop_k), int(lim), n, M, r,
                _lib.ptr(ii), _lib.ptr(jj),
                _lib.ptr(kk), E1, _lib.ptr(self._edge_src), _lib.ptr(self._edge_status), _lib.ptr(self._edge_tiles),
                self._edge_epoch, None, self.DIM, None, _lib.stream_ptr(dev)), "rvo_edges_step")
            # the rows of the hidden state follow on a side stream: the update graph waits for them (external event)
            # only before its first use, after reproject + corr
            ns = self._net_stream
            ns.wait_stream(cur)
            _lib.check(_lib.lib().rvo_net_rows(_lib.ptr(self.net), _lib.ptr(self._edge_src), E1, self.DIM, _lib.ptr(out),
                                               ctypes.c_void_p(ns.cuda_stream)), "rvo_net_rows")
            self._net_ready.record(ns)
            self._net_ready_fresh = True
        self.ii, self.jj, self.kk = ii, jj, kk
        self._net_swap(E1)
        self._plans = None

    def _graph_plans(self):
        if self._plans is None:
            self._plans = self._new_plans(self.ii, self.jj, self.kk)
        return self._plans

    @torch.no_grad()
    def motion_probe(self):
        """median |delta| of the newest patches against the candidate frame (Ramp_vo.py:210-225)"""
        kk = torch.arange(self.m - self.M, self.m, device=self.device)
        jj = self.n * torch.ones_like(kk)
        ii = self.ix[kk]
        net = torch.zeros(1, len(ii), self.DIM, device=self.device)
        coords = self.reproject(indicies=(ii, jj, kk))
        with torch.autocast("cuda", enabled=self.autocast):
            if self.autocast and self.P == 3:
                corr = self.corr_tiles(coords, indicies=(kk, jj))     # tcgen05 path, tile layout
            else:
                corr = self.corr(coords, indicies=(kk, jj))
            ctx = self.imap[:, kk % (self.M * self.mem)]
            net, (delta, weight, _) = self.network.update(net, ctx, corr, None, ii, jj, kk)
        return torch.quantile(delta.norm(dim=-1).float(), 0.5)

    def motionmag(self, i, j):
        """mean flow of the patches of frame i seen in frame j (Ramp_vo.py:227-235)"""
        k = (self.ii == i) & (self.jj == j)
        flow = pops.flow_mag(SE3(self.poses), self.patches, self.intrinsics, self.ii[k], self.jj[k],
                             self.kk[k], beta=0.5)
        return flow.mean().item()

    def sync(self, defer_removal=False):
        """finish a keyframe step deferred by pipeline mode (no-op otherwise).  defer_removal (internal, __call__ only):
        leave the removal of the edges that fell out of the window to the fused graph step of the new frame"""
        if self._pending_lim is not None:      # never observable from outside: flush a removal left pending
            lim, self._pending_lim = self._pending_lim, None
            self._flush_pending_drop()
            self.remove_factors(self.ii < lim, lambda i, j: i < lim)
        if self._pending_kf is not None:
            host, ev = self._pending_kf
            self._pending_kf = None
            ev.synchronize()
            vals = host.tolist()
            if vals[4] != 0.0:
                raise RuntimeError("rvo_edges_step: the device built %d edges, the host expected another count "
                                   "(patch-graph bookkeeping out of sync)" % (int(vals[4]) - 1))
            self._keyframe_finish(vals[:4], defer_removal=defer_removal)

    def _keyframe_begin(self):
        """launch the flow-magnitude reduction of the keyframe test (Ramp_vo.py:237-241); returns the device
        buffer [sum_ij, count_ij, sum_ji, count_ji]"""
        i = self.n - self.cfg.KEYFRAME_INDEX - 1
        j = self.n - self.cfg.KEYFRAME_INDEX + 1
        # motionmag(i, j) + motionmag(j, i) in one launch and one device->host read
        out4 = torch.empty(4, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().rvo_pair_flow(
                _lib.ptr(self.poses_), _lib.ptr(self.patches_), _lib.ptr(self.intrinsics_),
                _lib.ptr(self.ii), _lib.ptr(self.jj), _lib.ptr(self.kk), self.ii.numel(), self.P, i, j,
                0.5, _lib.ptr(out4), _lib.stream_ptr(self.device)), "rvo_pair_flow")
        if self.world_size > 1:     # edges i->j live on owner(i), j->i on owner(j): sum the partial flow sums
            import torch.distributed as dist
            dist.all_reduce(out4, op=dist.ReduceOp.SUM, group=self.group)
            self.collective_bytes += 16
            self.collective_calls += 1
        return out4

    def keyframe(self):
        """remove keyframe n-KEYFRAME_INDEX if motion is small (Ramp_vo.py:237-274)"""
        self.sync()
        self._keyframe_finish(self._keyframe_begin().tolist())

    def _keyframe_defer(self):
        out4 = self._keyframe_begin()
        host = getattr(self, "_kf_host", None)
        if host is None:
            host = self._kf_host = torch.zeros(5, dtype=torch.float32).pin_memory()
        host[:4].copy_(out4, non_blocking=True)
        if self._edge_status is not None:       # status of this frame's rvo_edges_step rides along
            host[4:].copy_(self._edge_status, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._pending_kf = (host, ev)
        self._frame_done = ev           # everything frame t put on the main stream has finished

    def _keyframe_finish(self, vals, defer_removal=False):
        s1, c1, s2, c2 = vals
        nan = float("nan")      # the reference takes the mean of an empty selection (nan) too
        m = (s1 / c1 if c1 else nan) + (s2 / c2 if c2 else nan)
        if m / 2 < self.cfg.KEYFRAME_THRESH:
            k = self.n - self.cfg.KEYFRAME_INDEX
            t0 = self.tstamps_[k - 1].item()
            t1 = self.tstamps_[k].item()
            dP = SE3(self.poses_[k]) * SE3(self.poses_[k - 1]).inv()
            self.delta[t1] = (t0, dP)
            if defer_removal:
                # the edge lists are left as they are: rvo_edges_step of the new frame removes the edges of k and
                # renumbers the rest in the same pass as the window removal and the appends (drop_k)
                pc = self._pair_counts()
                for key in [key for key in pc if key[0] == k or key[1] == k]:
                    del pc[key]
                self._pending_drop = k
            else:
                self.remove_factors((self.ii == k) | (self.jj == k), lambda i, j: i == k or j == k)
                # renumber without boolean indexing (no size-dependent sync): subtract masks
                gi = self.ii > k
                self.kk -= gi * self.M
                self.ii -= gi.to(self.ii.dtype)
                self.jj -= (self.jj > k).to(self.jj.dtype)
            self._pair_cnt = {(i - (i > k), j - (j > k)): v for (i, j), v in self._pair_cnt.items()}
            self._min_src = max(self._min_src - 1, 0)
            # shift every per-frame buffer one slot down (the reference loops frame by frame)
            n = self.n
            for buf in (self.tstamps_, self.colors_, self.poses_, self.patches_, self.intrinsics_):
                buf[k:n - 1] = buf[k + 1:n].clone()
            src = torch.arange(k + 1, n, device=self.device) % self.mem
            dst = torch.arange(k, n - 1, device=self.device) % self.mem
            if len(src):
                self.imap_[dst] = self.imap_[src]
                g = self._gmap_store.view(self.mem, self.M, self.P, self.P, 128)
                g[dst] = g[src]
                self._fmap1_store[dst] = self._fmap1_store[src]
                self._fmap2_store[dst] = self._fmap2_store[src]
            if len(self._owner) > k:
                del self._owner[k]
                if self._owner_dev is not None:
                    self._owner_dev[k:n - 1] = self._owner_dev[k + 1:n].clone()
            self.n -= 1
            self.m -= self.M
        lim = self.n - self.cfg.REMOVAL_WINDOW
        if defer_removal:           # __call__ fuses it with the new frame's appends (rvo_edges_step)
            self._pending_lim = lim
            return
        self.remove_factors(self.ii < lim, lambda i, j: i < lim)     # ix[kk] == ii (index_[f] == f)

    def _update_body(self, ii, jj, kk, net_in, net_out, plans, t0, t1, t0_dev=None, before_update=None,
                     with_ba=True):
        """reproject -> corr -> update operator -> 2 BA iterations on explicit buffers; every
        host-side scalar is either constant across frames or read from device memory (t0_dev).
        with_ba=False (sharded graphs): stops after the update operator and returns
        (net, (coords, delta, weight)) — BA then runs with its all-reduce outside the captured graph."""
        coords = self.reproject(indicies=(ii, jj, kk))
        with torch.autocast("cuda", enabled=self.autocast):
            if self.autocast and self.P == 3:
                corr = self.corr_tiles(coords, indicies=(kk, jj))     # tcgen05 path, tile layout
            else:
                corr = self.corr(coords, indicies=(kk, jj))
            ctx = (self.imap_, kk, self.M * self.mem)     # imap[:, kk % (M*mem)], gather fused
            if before_update is not None:
                before_update()                           # join the branch that built `plans`
            new_net, (delta, weight, _) = self.network.update(net_in, ctx, corr, None, ii, jj, kk,
                                                              plans=plans, net_out=net_out)
        if not with_ba:
            return new_net, (coords, delta.float().contiguous(), weight.float().contiguous())
        fused = (delta.dtype == torch.float32 and weight.dtype == torch.float32 and delta.is_contiguous()
                 and weight.is_contiguous() and coords.is_contiguous())
        try:
            if fused:
                # target formation + filter_features (Ramp_vo.py:288-296) happen inside the BA kernels' edge load
                if t0_dev is None:
                    t0_dev = torch.tensor([t0], dtype=torch.int32, device=self.device)
                    t1 = t1 - t0
                wf = torch.empty_like(weight)
                fastba.BA_fused(self.poses, self.patches, self.intrinsics, coords, delta, weight, self.ht // 4,
                                self.wd // 4, self.lmbda, ii, jj, plans.plan_k, t1, t0_dev, 2, weight_out=wf)
                weight = wf
            else:
                weight = weight.float()
                target = coords[..., self.P // 2, self.P // 2] + delta.float()
                weight = filter_features(confidences=weight, target=target, data_shape=(self.ht // 4, self.wd // 4))
                fastba.BA(self.poses, self.patches, self.intrinsics, target, weight, self.lmbda, ii, jj, kk,
                          t0, t1, M=self.M, iterations=2, eff_impl=False, plan=plans.plan_k, t0_dev=t0_dev)
        except RuntimeError as e:           # only a BA failure is non-fatal, like the reference (:302-306)
            if torch.cuda.is_current_stream_capturing():
                raise
            print(f"WARNING: BA failed...{e}")
        return new_net, weight

    def _new_plans(self, ii, jj, kk):
        # source frames that can still own edges: REMOVAL_WINDOW + 2 in steady state, and all 8 frames that exist
        # before initialisation (nothing is removed until n == 8, Ramp_vo.py:385-395) — the SoftAgg group buffers
        # are sized from this bound (a REMOVAL_WINDOW < 6 used to overflow them silently)
        frames = max(self.cfg.REMOVAL_WINDOW + 2, 8)
        return GraphPlans(ii, jj, kk, kmax=self.N * self.M, jmax=self.N, max_patches=frames * self.M,
                          max_pairs=frames * (2 * self.cfg.PATCH_LIFETIME + 1))

    @torch.no_grad()
    def update(self):
        """one recurrent update: reproject -> corr -> update operator -> 2 BA iterations
        (Ramp_vo.py:276-310).  With use_graphs the whole body (incl. the graph-plan sorts) is a CUDA
        graph keyed by (edge count, window length), replayed while the window slides."""
        self.sync()
        E = self.ii.numel()
        t0 = self.n - self.cfg.OPTIMIZATION_WINDOW if self.is_initialized else 1
        t0 = max(t0, 1)
        if self.world_size > 1:
            return self._update_sharded(E, t0)
        key = (E, self.n - t0, self._net_cur)
        repeat = key == self._ukey_prev or key in self._ukey_hist   # capture only shapes that recur
        self._ukey_hist = (self._ukey_hist + [self._ukey_prev])[-4:]
        self._ukey_prev = key
        if (self.use_graphs and self.autocast and E > 0 and self.network.update._fused_ready()
                and (repeat or (key + (self._net_bufs[0].data_ptr(),)) in self._ugraphs)):
            self._update_graphed(E, t0, self.n)
        else:
            self._net_join()
            plans = self._graph_plans()
            other = self._net_other(E)
            new_net, weight = self._update_body(self.ii, self.jj, self.kk, self.net, other, plans, t0, self.n)
            if new_net.data_ptr() != other.data_ptr():
                other.copy_(new_net)   # generic (non-fused) path returned its own tensor
            self._net_swap(E)
            self.last_weight = weight
        pts = pops.point_cloud_centers(SE3(self.poses), self.patches[:, :self.m], self.intrinsics,
                                       self.ix[:self.m])
        self.points_[:len(pts)] = pts

    def _update_sharded(self, E, t0):
        """one recurrent update over THIS rank's edges: reproject / corr / update operator
        are local (a CUDA graph once the shape recurs); BA = local assembly -> all-reduce of [S | y] -> identical
        solve on every rank, twice; then the refined depths of the active frames are exchanged."""
        from . import sharded
        key = (E, self.n - t0, self._net_cur)
        repeat = key == self._ukey_prev or key in self._ukey_hist
        self._ukey_hist = (self._ukey_hist + [self._ukey_prev])[-4:]
        self._ukey_prev = key
        P = self.P
        plans = None
        if E == 0:      # this rank owns no edge yet: it still takes part in the collectives
            z = torch.zeros(1, 0, 2, device=self.device)
            coords, delta, weight = torch.zeros(1, 0, 2, P, P, device=self.device), z, z
        elif (self.use_graphs and self.autocast and self.network.update._fused_ready()
              and (repeat or (key + (self._net_bufs[0].data_ptr(),)) in self._ugraphs)):
            gkey = key + (self._net_bufs[0].data_ptr(),)
            g = self._ugraphs.get(gkey)
            if g is None:
                if len(self._ugraphs) >= 8:
                    self._ugraphs.pop(next(iter(self._ugraphs)))
                g = self._ugraphs[gkey] = _UpdateGraph(self, E, self.n - t0)
            g.run(t0)
            self._net_swap(E)
            coords, delta, weight = g.weight
            plans = g.plans
        else:
            plans = self._graph_plans()
            other = self._net_other(E)
            new_net, (coords, delta, weight) = self._update_body(self.ii, self.jj, self.kk, self.net, other, plans,
                                                                 t0, self.n, with_ba=False)
            if new_net.data_ptr() != other.data_ptr():
                other.copy_(new_net)
            self._net_swap(E)
        wf = torch.empty_like(weight)
        n6 = 6 * (self.n - t0)
        try:
            sharded.sharded_BA_fused(self.poses, self.patches, self.intrinsics, coords, delta, weight, self.ht // 4,
                                     self.wd // 4, self.lmbda, self.ii, self.jj, self.kk, t0, self.n, iterations=2,
                                     weight_out=wf, group=self.group,
                                     plan=plans.plan_k if plans is not None else None)
        except RuntimeError as e:
            print(f"WARNING: BA failed...{e}")
        self.last_weight = wf
        self.collective_bytes += 2 * n6 * (n6 + 1) * 4
        self.collective_calls += 2
        lo = max(self.n - self.cfg.REMOVAL_WINDOW - 2, 0)
        sharded.exchange_depths_owned(self.patches_, self._owner_dev if self._owner_dev is not None else self._owner,
                                      lo, self.n, self.rank, self.group)
        self.collective_bytes += (self.n - lo) * self.M * P * P * 4
        self.collective_calls += 1
        pts = pops.point_cloud_centers(SE3(self.poses), self.patches[:, :self.m], self.intrinsics, self.ix[:self.m])
        self.points_[:len(pts)] = pts

    def _update_graphed(self, E, t0, t1):
        key = (E, t1 - t0, self._net_cur, self._net_bufs[0].data_ptr())
        g = self._ugraphs.get(key)
        if g is None:
            if len(self._ugraphs) >= 8:              # bounded cache: drop the oldest capture
                self._ugraphs.pop(next(iter(self._ugraphs)))
            g = self._ugraphs[key] = _UpdateGraph(self, E, t1 - t0)
        g.run(t0)
        self._last_ugraph = g
        self._net_swap(E)
        self.last_weight = g.weight

    def _edges_forw(self):
        """patches of frames [n-r, n-1) -> frame n-1 (Ramp_vo.py:312-318): (kk, jj, pair counts)"""
        r = self.cfg.PATCH_LIFETIME
        f0, f1 = max(self.n - r, 0), max(self.n - 1, 0)
        if self.world_size > 1:     # only the patches of the frames this rank owns
            fr = [f for f in range(f0, f1) if self._owner[f] == self.rank]
            base = torch.tensor(fr, dtype=torch.long, device=self.device) * self.M
            kk = (base[:, None] + torch.arange(self.M, device=self.device)[None]).reshape(-1)
            return kk, torch.full_like(kk, self.n - 1), {(i, self.n - 1): self.M for i in fr}
        kk = torch.arange(self.M * f0, self.M * f1, device=self.device)
        return kk, torch.full_like(kk, self.n - 1), {(i, self.n - 1): self.M for i in range(f0, f1)}

    def _edges_back(self):
        """patches of frame n-1 -> frames [n-r, n) (Ramp_vo.py:320-325), 'ij' meshgrid order"""
        r = self.cfg.PATCH_LIFETIME
        t0 = self.M * max((self.n - 1), 0)
        t1 = self.M * max((self.n - 0), 0)
        j0 = max(self.n - r, 0)
        if self.world_size > 1 and self.n > 0 and self._owner[self.n - 1] != self.rank:
            e = torch.zeros(0, dtype=torch.long, device=self.device)
            return e, e.clone(), {}
        k = torch.arange(t0, t1, device=self.device)
        j = torch.arange(j0, self.n, device=self.device)
        pairs = {(self.n - 1, jf): t1 - t0 for jf in range(j0, self.n)} if t1 > t0 else {}
        return k.repeat_interleave(self.n - j0), j.repeat(t1 - t0), pairs

    @torch.no_grad()
    def __call__(self, tstamp, input_tensor, intrinsics):
        """track a new frame (Ramp_vo.py:327-410)"""
        input_ = preprocess_input(input_tensor=input_tensor)
        P, M = self.P, self.M
        events, images, mask = input_
        mask_l = torch.as_tensor(mask).reshape(-1).tolist()
        graphable = (self.use_graphs and events.shape[1] == 1 and images.shape[1] == 1 and mask_l == [True]
                     and tuple(events.shape[-2:]) == (self.ht, self.wd)
                     and self.network.input_mode == "MultiScale")    # the SingleScale encoder runs eagerly
        if graphable:
            if self._pgraph is None:
                self.sync()
                self._pgraph = _PatchifyGraph(self)
            g = self._pgraph
            if self.encoder_after_corr and self._last_ugraph is not None:
                g.enc_stream.wait_event(self._last_ugraph.mid_event)
            if self.encoder_in_gap and getattr(self, "_frame_done", None) is not None:
                # the recurrent chain update(t-1) -> keyframe decision -> host bookkeeping -> update(t) leaves the GPU
                # idle while the host does its bookkeeping (~0.5 ms); start this frame's encoder exactly then,
                # instead of letting it compete with update(t-1) for the SMs
                g.enc_stream.wait_event(self._frame_done)
            g.run(events, images, reinit=(tstamp == 0))      # writes only the graph's staging buffers
        # pipeline mode: the previous frame's keyframe step, overlapped with the encoder graph.  In the steady state
        # the removal of the edges that left the window is fused with this frame's appends (rvo_edges_step)
        fast = (graphable and self.fast_edges and self.is_initialized and self.world_size == 1 and self.pipeline)
        # host work that does not depend on the keyframe decision goes BEFORE the wait for it
        intr = torch.as_tensor(intrinsics, dtype=torch.float32).reshape(-1).tolist() \
            if not (torch.is_tensor(intrinsics) and intrinsics.is_cuda) else intrinsics.float().cpu().tolist()
        intr4 = (ctypes.c_float * 4)(*[v / self.RES for v in intr])
        self.sync(defer_removal=fast)
        slot = self.n % self.mem
        gslot_store = self._gmap_store[slot * M:(slot + 1) * M]
        if graphable:
            torch.cuda.current_stream(self.device).wait_event(g.replayed)     # join the encoder stream
            patches, clr = g.patches, g.clr
            ring = [(g.gmap, gslot_store), (g.imap, self.imap_[slot]), (g.f1, self._fmap1_store[slot]),
                    (g.f2, self._fmap2_store[slot])]
        else:
            if not (events.is_cuda and images.is_cuda):
                input_ = (events.to(self.device), images.to(self.device), mask)
            gslot = gslot_store.permute(0, 3, 1, 2)[None]                                  # [1,M,128,P,P]
            with torch.autocast("cuda", enabled=self.autocast):
                fmap, gmap, imap, patches, _, clr = self.network.patchify(
                    input_=input_, patches_per_image=M, event_bias=self.event_bias,
                    reinit_hidden=True if tstamp == 0 else False, gmap_out=gslot)
            if fmap is None:
                return      # events only: the super state was updated, the VO is not
            f1_new = fmap[0, 0].permute(1, 2, 0).to(self.fdtype).contiguous()             # [h,w,128]
            ring = [(imap.reshape(M, self.DIM).to(self.fdtype).contiguous(), self.imap_[slot]),
                    (f1_new, self._fmap1_store[slot]), (pyramid_level2(f1_new), self._fmap2_store[slot])]

        # state writes of the new frame (Ramp_vo.py:345-372) in one launch: tstamps / intrinsics / index rows,
        # colours, depth initialisation (uniform draw, or the median of the last 3 frames) and patches_[n]
        self.tlist.append(tstamp)
        rnd = None if self.is_initialized else torch.rand_like(patches[:, :, 2, 0, 0, None, None]).contiguous()
        pn = patches if patches.dtype == torch.float32 and patches.is_contiguous() else patches.float().contiguous()
        cl = clr if clr.dtype == torch.float32 and clr.is_contiguous() else clr.float().contiguous()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().rvo_frame_commit(
                _lib.ptr(pn), _lib.ptr(cl), _lib.ptr(rnd), _lib.ptr(self.patches_), _lib.ptr(self.tstamps_),
                _lib.ptr(self.intrinsics_), _lib.ptr(self.index_), _lib.ptr(self.index_map_), _lib.ptr(self.colors_),
                intr4, self.n, M, P, self.N, self.counter, self.m + M, 3 if self.is_initialized else 0,
                _lib.stream_ptr(self.device)), "rvo_frame_commit")

        if self.n > 1:
            if self.cfg.MOTION_MODEL == 'DAMPED_LINEAR':
                with torch.cuda.device(self.device):
                    _lib.check(_lib.lib().rvo_motion_model(_lib.ptr(self.poses_), self.n,
                                                           float(self.cfg.MOTION_DAMPING),
                                                           _lib.stream_ptr(self.device)), "rvo_motion_model")
            else:
                self.poses_[self.n] = self.poses_[self.n - 1]

        # network attributes: gmap / imap / fmap pyramid into their ring slots (Ramp_vo.py:376-381), one launch
        copy_segments(ring)
        if graphable:
            g.consumed.record(torch.cuda.current_stream(self.device))        # staging buffers may be overwritten

        self.counter += 1
        if self.n > 0 and not self.is_initialized:
            if self.motion_probe() < 2.0:
                self.delta[self.counter - 1] = (self.counter - 2, self.Id[0])
                return

        self._owner = self._owner[:self.n] + [(self.counter - 1) % self.world_size]
        if self.world_size > 1:
            if self._owner_dev is None:
                self._owner_dev = torch.full((self.N,), -1, dtype=torch.int32, device=self.device)
            self._owner_dev[self.n] = self._owner[-1]
        self.n += 1
        self.m += self.M
        if self._pending_lim is not None:
            lim, self._pending_lim = self._pending_lim, None
            self._edges_step(lim)
        else:
            self.append_factors(*self._edges_forw())
            self.append_factors(*self._edges_back())

        if self.n == 8 and not self.is_initialized:
            self.is_initialized = True
            for _ in range(12):
                self.update()
        elif self.is_initialized:
            self.update()
            if self.pipeline:
                self._keyframe_defer()
            else:
                self.keyframe()
