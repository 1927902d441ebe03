"""Python owner of a dpb200 model + sessions (device memory via torch, compute via libdpb200.so)."""
import ctypes as C
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib
from ._lib import lib, check
from .config import EXTRA_HEAD_ORDER, ModelSpec
from .weights import Packed, pack_state_dict

_DTYPES = {0: torch.bfloat16, 1: torch.float32, 2: torch.int32, 3: torch.uint8}


class Session:
    """Launch plan for one (batch, H0, W0, input dtype, output dtype) on one device. `out_half`: the four
    DensePose tensors are produced as fp16 by the kernel itself (the reference's `.half()` contract)."""

    def __init__(self, engine: "Engine", batch: int, h0: int, w0: int, src_u8: bool, out_half: bool = False):
        self.engine, self.batch, self.h0, self.w0, self.src_u8 = engine, batch, h0, w0, src_u8
        self.out_half = out_half
        spec = engine.spec
        nbytes = lib.dpb200_session_workspace_bytes(engine.handle, batch, h0, w0)
        if nbytes == 0:
            raise _lib.DPB200Error("session_workspace_bytes failed: " + _lib.last_error())
        self.workspace = torch.empty(nbytes, dtype=torch.uint8, device=engine.device)
        self.workspace.zero_()
        h = C.c_void_p()
        check(lib.dpb200_session_create(engine.handle, batch, h0, w0, int(src_u8), self.workspace.data_ptr(),
                                        nbytes, C.byref(h)), "dpb200_session_create")
        self.handle = h
        n = batch * spec.dets_per_image
        s = spec.out_size
        dev = engine.device
        self.pred_boxes = torch.zeros(batch, spec.dets_per_image, 4, device=dev)
        self.scores = torch.zeros(batch, spec.dets_per_image, device=dev)
        self.det_count = torch.zeros(batch, dtype=torch.int32, device=dev)
        self.det_offsets = torch.zeros(batch + 1, dtype=torch.int32, device=dev)
        odt = torch.float16 if out_half else torch.float32
        self.coarse = torch.zeros(n, spec.coarse_ch, s, s, device=dev, dtype=odt)
        self.fine = torch.zeros(n, 25, s, s, device=dev, dtype=odt)
        self.u = torch.zeros(n, 25, s, s, device=dev, dtype=odt)
        self.v = torch.zeros(n, 25, s, s, device=dev, dtype=odt)
        # confidence heads of a WC* model (sigma_2 / kappa_u / kappa_v / segm confidences): extra outputs (f4)
        self.extra = {name: torch.zeros(n, ch, s, s, device=dev, dtype=odt) for name, ch in spec.extra_heads}
        self.io = _lib.ForwardIO()
        self.io.pred_boxes = self.pred_boxes.data_ptr(); self.io.scores = self.scores.data_ptr()
        self.io.det_count = self.det_count.data_ptr(); self.io.det_offsets = self.det_offsets.data_ptr()
        self.io.coarse = self.coarse.data_ptr(); self.io.fine = self.fine.data_ptr()
        self.io.u = self.u.data_ptr(); self.io.v = self.v.data_ptr()
        self.io.out_half = int(out_half)
        for i, name in enumerate(EXTRA_HEAD_ORDER):
            self.io.extra[i] = self.extra[name].data_ptr() if name in self.extra else None
        # launches go to a private stream (the legacy default stream cannot be graph-captured); run() orders
        # it after / before the caller's current stream with events, so the semantics stay "enqueued on the
        # current stream"
        self.stream = torch.cuda.Stream(device=dev)
        self.use_graph = engine.use_graph
        check(lib.dpb200_session_set_graph(self.handle, int(self.use_graph)), "dpb200_session_set_graph")
        g = (C.c_int32 * 4)()
        lib.dpb200_session_geometry(self.handle, C.byref(g))
        self.hr, self.wr, self.hp, self.wp = g[0], g[1], g[2], g[3]

    def __del__(self):
        h = getattr(self, "handle", None)
        if h:
            lib.dpb200_session_destroy(h)
            self.handle = None

    @property
    def launches(self) -> int:
        return lib.dpb200_session_launch_count(self.handle)

    def run(self, images: torch.Tensor, bgr: bool = True):
        """images: [B,H0,W0,3] contiguous on the engine's device (fp32, or uint8 for a u8 session).
        Enqueues the whole forward on the current stream; results land in the session's output tensors."""
        assert images.is_cuda and images.is_contiguous() and tuple(images.shape) == (self.batch, self.h0, self.w0, 3)
        assert images.dtype == (torch.uint8 if self.src_u8 else torch.float32)
        self.io.images = images.data_ptr()
        self.io.bgr = int(bgr)
        cur = torch.cuda.current_stream()
        if not self.use_graph:
            check(lib.dpb200_session_run(self.handle, C.byref(self.io), C.c_void_p(cur.cuda_stream)), "dpb200_session_run")
            return
        self.stream.wait_stream(cur)
        check(lib.dpb200_session_run(self.handle, C.byref(self.io), C.c_void_p(self.stream.cuda_stream)),
              "dpb200_session_run")
        cur.wait_stream(self.stream)

    def op_info(self) -> List[Tuple[str, float]]:
        out = []
        buf = C.create_string_buffer(160)
        fl = C.c_double()
        for i in range(self.launches):
            check(lib.dpb200_session_op_info(self.handle, i, buf, 160, C.byref(fl)), "op_info")
            out.append((buf.value.decode(), fl.value))
        return out

    def op_bytes(self) -> List[float]:
        """Algorithmic HBM bytes per launch (full detection capacity)."""
        out = []
        by = C.c_double()
        for i in range(self.launches):
            check(lib.dpb200_session_op_bytes(self.handle, i, C.byref(by)), "op_bytes")
            out.append(by.value)
        return out

    def profile(self, images: torch.Tensor, bgr: bool = True) -> List[float]:
        """One run with CUDA events around every launch (on the current stream); returns ms per launch."""
        self.io.images = images.data_ptr()
        self.io.bgr = int(bgr)
        n = self.launches
        ms = (C.c_float * n)()
        check(lib.dpb200_session_profile(self.handle, C.byref(self.io),
                                         C.c_void_p(torch.cuda.current_stream().cuda_stream), ms, n), "session_profile")
        return list(ms)

    def tap(self, name: str) -> torch.Tensor:
        """View of an intermediate tensor inside the workspace (stage-parity tests)."""
        p = C.c_void_p(); shape = (C.c_int64 * 4)(); dt = C.c_int32()
        check(lib.dpb200_session_tap(self.handle, name.encode(), C.byref(p), C.byref(shape), C.byref(dt)), "tap")
        dtype = _DTYPES[dt.value]
        off = p.value - self.workspace.data_ptr()
        numel = shape[0] * shape[1] * shape[2] * shape[3]
        nb = numel * torch.empty(0, dtype=dtype).element_size()
        return self.workspace[off:off + nb].view(dtype).view(shape[0], shape[1], shape[2], shape[3])

    def results(self, copy: bool = True) -> List[Dict[str, torch.Tensor]]:
        """Per-image result dicts in the reference's output format (synchronises). copy=True (default) returns fresh
        tensors; copy=False returns VIEWS of this session's persistent output buffers, which the next run() of the
        session overwrites in place (zero-copy fast path for callers that consume the results first)."""
        counts = self.det_count.cpu().tolist()
        offs = self.det_offsets.cpu().tolist()
        own = (lambda t: t.clone()) if copy else (lambda t: t)
        out = []
        for b in range(self.batch):
            d, o = counts[b], offs[b]
            out.append({
                "image_size": torch.tensor([self.h0, self.w0], dtype=torch.int64, device=self.engine.device),
                "pred_boxes": own(self.pred_boxes[b, :d]),
                "scores": own(self.scores[b, :d]),
                "pred_classes": torch.zeros(d, dtype=torch.int64, device=self.engine.device),
                "pred_densepose_coarse_segm": own(self.coarse[o:o + d]),
                "pred_densepose_fine_segm": own(self.fine[o:o + d]),
                "pred_densepose_u": own(self.u[o:o + d]),
                "pred_densepose_v": own(self.v[o:o + d]),
            })
            for name, t in self.extra.items():      # WC* models: the confidence heads the reference builds but never emits
                out[-1]["pred_densepose_" + name] = own(t[o:o + d])
        return out


class Engine:
    """Packed weights + native model handle on one CUDA device."""

    def __init__(self, spec: ModelSpec, state_dict: Optional[Dict[str, torch.Tensor]] = None,
                 packed: Optional[Dict[str, Packed]] = None, device: Optional[torch.device] = None,
                 use_graph: bool = True, max_sessions: int = 4, strict: bool = False, resize_variant: int = 0):
        """strict: fp32-class numerics end to end (activations and weights as bf16 hi/lo pairs, three tensor-core passes
        per product, fp32 accumulate; `packed` weights must then come from pack_state_dict(..., strict=True)) — the mode
        in which proposals, NMS keep lists, detection counts and label maps are compared with the reference by index.
        max_sessions: how many plain (slot 0) sessions of distinct shapes stay cached; each owns a workspace
        (0.9 GB at batch 1, 7.1 GB at batch 8 for R50-s1x), so a stream of differently sized images must not
        accumulate them. The least recently used one is dropped; HostPipeline slots are released by close()."""
        _lib.require_device()
        self.spec = spec
        self.strict = strict
        self.max_sessions = max(1, max_sessions)
        self.use_graph = use_graph
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        if packed is None:
            if state_dict is None:
                raise ValueError("Engine needs a state_dict or packed weights")
            packed = pack_state_dict(state_dict, spec, self.device, strict=strict)
        self.packed = packed
        cfg = _lib.ModelConfig()
        cfg.depth = spec.depth; cfg.head = 0 if spec.head == "v1convx" else 1
        cfg.decoder_on = int(spec.decoder_on); cfg.pooler_res = spec.pooler_res; cfg.coarse_ch = spec.coarse_ch
        cfg.score_thresh = spec.score_thresh; cfg.nms_test = spec.nms_test; cfg.rpn_nms = spec.rpn_nms
        cfg.dets_per_image = spec.dets_per_image; cfg.rpn_pre_topk = spec.rpn_pre_topk
        cfg.rpn_post_topk = spec.rpn_post_topk; cfg.min_size = spec.min_size; cfg.max_size = spec.max_size
        for i in range(3):
            cfg.pixel_mean[i] = spec.pixel_mean[i]; cfg.pixel_std[i] = spec.pixel_std[i]
        cfg.input_rgb = int(spec.input_format == "RGB")
        cfg.strict = int(strict)
        cfg.resize_variant = int(resize_variant)     # 0: ATen's multi-threaded float resize kernel, 1: its single-threaded one
        heads = dict(spec.extra_heads)
        for i, name in enumerate(EXTRA_HEAD_ORDER):
            cfg.extra_ch[i] = heads.get(name, 0)
        arr = (_lib.Weight * len(packed))()
        self._names = []
        for i, (name, (d0, d1, cin_pad, cout_pad)) in enumerate(packed.items()):
            nb = name.encode()
            self._names.append(nb)
            arr[i].name = nb
            arr[i].data0 = d0.data_ptr()
            arr[i].data1 = d1.data_ptr() if d1 is not None else None
            arr[i].cin_pad, arr[i].cout_pad = cin_pad, cout_pad
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(lib.dpb200_model_create(C.byref(cfg), arr, len(packed), C.byref(h)), "dpb200_model_create")
        self.handle = h
        self._sessions: Dict[Tuple[int, int, int, bool, int, bool], Session] = {}

    def __del__(self):
        h = getattr(self, "handle", None)
        if h:
            self._sessions.clear()
            lib.dpb200_model_destroy(h)
            self.handle = None

    def session(self, batch: int, h0: int, w0: int, src_u8: bool = False, slot: int = 0,
                out_half: bool = False) -> Session:
        """`slot` > 0 gives additional independent sessions (own workspace, outputs, stream) of the same shape."""
        key = (batch, h0, w0, src_u8, slot, out_half)
        s = self._sessions.pop(key, None)
        if s is None:
            plain = [k for k in self._sessions if k[4] == 0]
            if slot == 0 and len(plain) >= self.max_sessions:
                del self._sessions[plain[0]]                      # dicts keep insertion order: [0] is the LRU
            with torch.cuda.device(self.device):
                s = Session(self, batch, h0, w0, src_u8, out_half)
        self._sessions[key] = s                                   # (re)insert as most recently used
        return s

    def forward_batch(self, images: torch.Tensor, bgr: bool = True, out_half: bool = False,
                      copy: bool = True) -> List[Dict[str, torch.Tensor]]:
        """images [B,H,W,3] (HWC, fp32 or uint8) on any device -> list of reference-format result dicts.
        copy=False returns views of the session's output buffers (see Session.results): they are overwritten by the
        next forward_batch of the same (batch, H, W, dtype)."""
        images = images.to(self.device, non_blocking=True).contiguous()
        if images.dtype not in (torch.uint8, torch.float32):
            images = images.float()
        s = self.session(images.shape[0], images.shape[1], images.shape[2], images.dtype == torch.uint8,
                         out_half=out_half)
        with torch.cuda.device(self.device):
            s.run(images, bgr)
        return s.results(copy=copy)


class HostPipeline:
    """Host-in / host-out serving loop: `depth` sessions of one shape, each with its own stream and pinned
    staging buffers, so the PCIe copies of one batch overlap the kernels of the next and the host never blocks on a
    copy it has just issued.

        pipe = HostPipeline(engine, batch, h0, w0)
        for frames in batches:            # frames: [B,H,W,3] host tensor (pinned or not)
            done = pipe.submit(frames)    # results of the batch submitted `depth + 1` calls ago (or None)
        tail = pipe.drain()               # everything still in flight, oldest first

    submit(k) does three things, in this order:
      1. the slot it is about to reuse ran batch k - depth: its forward has finished, its boxes / scores / counts are on
         the host, so the DensePose tensors are copied back COUNT-AWARE — the rows are packed by `det_offsets` on the
         device, only the `dp_total` rows that hold detections cross PCIe, not the B x dets_per_image capacity (386 MB
         per image at 100 detections, 3.9 MB per detection). The copies are enqueued on the slot's stream, not waited for;
      2. batch k is enqueued behind them (H2D copy -> forward -> D2H of boxes / scores / counts);
      3. the copies issued by the PREVIOUS call have had a whole call to finish: they are waited for and returned.

    One pipeline belongs to one host thread (submit / drain are not re-entrant); use one pipeline per thread or GPU.
    Lifetime of returned tensors: they are views of pinned host buffers owned by the pipeline; `depth + 1` result sets
    rotate, so what one submit() returns stays valid until the NEXT submit() has returned, and everything drain()
    returns is valid together. Copy what must live longer.

    extract=True is the run.py flow (run.py:33-57 + visualizer.py:46-56) as a pipeline: only boxes / scores /
    counts come back after the forward; in step 1 the per-box resample + part argmax + U/V gather
    (`dpb200_dp_resample`) runs on the device over that batch's detections and only `labels` (uint8, or int64
    like the reference with labels_u8=False) and `uv` at box resolution cross PCIe. Each result dict then holds
    pred_boxes, scores, boxes_xywh and `densepose` = [{'labels': [h,w], 'uv': [2,h,w]} per detection].
    """

    def __init__(self, engine: Engine, batch: int, h0: int, w0: int, src_u8: bool = False, depth: int = 2,
                 out_half: bool = False, extract: bool = False, labels_u8: bool = True):
        self.engine, self.depth = engine, depth
        self.extract, self.labels_u8 = extract, labels_u8
        if extract and out_half:
            raise ValueError("extract=True reads the fp32 DensePose tensors on the device; out_half is for full outputs")
        self.extract_d2h_bytes = 0          # bytes of the last returned extraction (data dependent)
        self.last_d2h_bytes = 0             # bytes the last RETURNED batch moved device -> host (count-aware)
        self.slots = []
        dt = torch.uint8 if src_u8 else torch.float32
        with torch.cuda.device(engine.device):
            for i in range(depth):
                sess = engine.session(batch, h0, w0, src_u8, slot=i + 1, out_half=out_half)
                dev_in = torch.empty(batch, h0, w0, 3, dtype=dt, device=engine.device)
                host_in = torch.empty(batch, h0, w0, 3, dtype=dt).pin_memory()
                small_dev = [sess.pred_boxes, sess.scores, sess.det_count, sess.det_offsets]
                small_host = [torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in small_dev]
                self.slots.append(dict(sess=sess, dev_in=dev_in, host_in=host_in, small_dev=small_dev,
                                       small_host=small_host, done=torch.cuda.Event(), busy=False,
                                       ex_dev=[None, None], ex_host=[None, None], ex_flip=0))
                if extract:
                    # both extraction buffer sets of the slot exist before the first batch: a pinned allocation
                    # (cudaHostAlloc, milliseconds) never lands inside a serving loop unless a batch outgrows them
                    for flip in (0, 1):
                        self._ex_buffers(self.slots[-1], flip, 0)
            # depth + 1 rotating result sets for the big tensors (full capacity, pinned): up to depth batches in the
            # slots plus one whose copy has been issued can be in flight, and a returned set is not refilled before the
            # next submit() has returned
            s0 = self.slots[0]["sess"]
            self.big_dev_names = ("coarse", "fine", "u", "v") + tuple(n for n, _ in engine.spec.extra_heads)
            self.sets = []
            if not extract:
                for _ in range(depth + 1):
                    self.sets.append([torch.empty(self._big(s0, n).shape, dtype=self._big(s0, n).dtype).pin_memory()
                                      for n in self.big_dev_names])
        self._set = 0
        self._pending = None                # ticket of the batch whose D2H has been issued but not yet returned
        self.h2d_bytes = self.slots[0]["host_in"].numel() * self.slots[0]["host_in"].element_size()
        self.small_d2h_bytes = sum(t.numel() * t.element_size() for t in self.slots[0]["small_host"])
        self.row_bytes = 0 if extract else sum(t[0].numel() * t.element_size() for t in self.sets[0])
        # capacity figure (every row): what a count-unaware copy would move per step
        self.d2h_bytes = self.small_d2h_bytes + self.row_bytes * batch * engine.spec.dets_per_image
        self._next = 0

    @staticmethod
    def _big(sess, name):
        return sess.extra[name] if name in sess.extra else getattr(sess, name)

    def close(self):
        """Drops the slots' sessions (workspaces, output and staging buffers) from the engine."""
        for sl in self.slots:
            sess = sl["sess"]
            for k, v in list(self.engine._sessions.items()):
                if v is sess:
                    del self.engine._sessions[k]
        self.slots = []
        self.sets = []
        self._pending = None
        # hand the pinned result sets (GBs at full capacity) back to the OS instead of leaving them in torch's pinned
        # cache: a process that opens pipelines of several shapes / dtypes must not accumulate them
        empty = getattr(torch._C, "_host_emptyCache", None)
        if empty is not None:
            empty()

    # ---- step 1: the slot's forward is done -> enqueue the (count-aware) copies of its results, do not wait
    def _issue(self, sl) -> Optional[dict]:
        if not sl["busy"]:
            return None
        sl["done"].synchronize()            # forward + small copies, enqueued `depth` calls ago
        sl["busy"] = False
        # boxes / scores / counts leave the slot's staging buffers (the slot is re-enqueued right after this)
        boxes, scores, counts, offs = [t.clone() for t in sl["small_host"]]
        sess = sl["sess"]
        ticket = dict(sess=sess, boxes=boxes, scores=scores, counts=counts, offs=offs, event=torch.cuda.Event())
        if self.extract:
            self._issue_extraction(sl, ticket)
        else:
            n = int(offs[sess.batch])                      # dp_total: rows that hold detections
            host = self.sets[self._set]
            self._set = (self._set + 1) % len(self.sets)
            if n:
                with torch.cuda.device(self.engine.device), torch.cuda.stream(sess.stream):
                    for hbuf, name in zip(host, self.big_dev_names):
                        hbuf[:n].copy_(self._big(sess, name)[:n], non_blocking=True)
            ticket.update(host=host, bytes=self.small_d2h_bytes + n * self.row_bytes)
        ticket["event"].record(sess.stream)
        return ticket

    def _ex_buffers(self, sl, flip: int, total: int):
        """Device + pinned host buffers of one extraction (labels, uv) holding at least `total` pixels."""
        ex_dev, ex_host = sl["ex_dev"][flip], sl["ex_host"][flip]
        if ex_dev is None or ex_dev[0].numel() < total:
            dev = self.engine.device
            lab_dt = torch.uint8 if self.labels_u8 else torch.int64
            cap = max(int(total * 1.25), 1 << 20)
            with torch.cuda.device(dev):
                ex_dev = (torch.empty(cap, dtype=lab_dt, device=dev), torch.empty(2 * cap, dtype=torch.float32, device=dev))
            ex_host = (torch.empty(cap, dtype=lab_dt).pin_memory(), torch.empty(2 * cap, dtype=torch.float32).pin_memory())
            sl["ex_dev"][flip], sl["ex_host"][flip] = ex_dev, ex_host
        return ex_dev, ex_host

    def _issue_extraction(self, sl, ticket):
        from . import ops
        sess, boxes, counts = ticket["sess"], ticket["boxes"], ticket["counts"]
        dev = self.engine.device
        cnt = [int(c) for c in counts]
        # packed detection order = the packed DensePose rows (det_offsets)
        packed = torch.cat([boxes[b, :cnt[b]] for b in range(sess.batch)]) if sum(cnt) else boxes.new_zeros((0, 4))
        boxes_xywh, wh, offsets = ops.box_sizes(packed)
        total = int(offsets[-1])
        lab_b = 1 if self.labels_u8 else 8
        flip = sl["ex_flip"]                               # two buffer sets per slot: what was returned stays intact
        sl["ex_flip"] ^= 1
        ex_dev, ex_host = self._ex_buffers(sl, flip, total)
        n = int(sum(cnt))
        if n and total:
            with torch.cuda.device(dev), torch.cuda.stream(sess.stream):
                wh_d = wh.to(dev, non_blocking=True)
                off_d = offsets.to(dev, non_blocking=True)
                ops.dp_resample_into(sess.coarse[:n], sess.fine[:n], sess.u[:n], sess.v[:n], wh_d, off_d, total,
                                     ex_dev[0], ex_dev[1], stream=C.c_void_p(sess.stream.cuda_stream))
                ex_host[0][:total].copy_(ex_dev[0][:total], non_blocking=True)
                ex_host[1][:2 * total].copy_(ex_dev[1][:2 * total], non_blocking=True)
            ticket["keep"] = (wh_d, off_d)                # alive until the kernel has run
        ticket.update(cnt=cnt, boxes_xywh=boxes_xywh, wh=wh, offsets=offsets, ex_host=ex_host,
                      bytes=self.small_d2h_bytes + total * (lab_b + 8), ex_bytes=total * (lab_b + 8))

    # ---- step 3: wait for copies issued one call ago, build the reference-format result dicts over the pinned buffers
    def _finish(self, ticket) -> Optional[List[Dict[str, object]]]:
        if ticket is None:
            return None
        ticket["event"].synchronize()
        sess, boxes, scores = ticket["sess"], ticket["boxes"], ticket["scores"]
        self.last_d2h_bytes = ticket["bytes"]
        out = []
        if self.extract:
            self.extract_d2h_bytes = ticket["ex_bytes"]
            cnt, boxes_xywh = ticket["cnt"], ticket["boxes_xywh"]
            # plain Python ints for the per-detection views (indexing a tensor per detection costs more than the kernel)
            whl, offl = ticket["wh"].tolist(), ticket["offsets"].tolist()
            labels_h, uv_h = ticket["ex_host"]
            k = 0
            for b in range(sess.batch):
                dens = []
                for _ in range(cnt[b]):
                    (w, h), o = whl[k], offl[k]
                    dens.append({"labels": labels_h[o:o + h * w].view(h, w), "uv": uv_h[2 * o:2 * o + 2 * h * w].view(2, h, w)})
                    k += 1
                out.append({"image_size": torch.tensor([sess.h0, sess.w0], dtype=torch.int64),
                            "pred_boxes": boxes[b, :cnt[b]], "scores": scores[b, :cnt[b]],
                            "boxes_xywh": boxes_xywh[k - cnt[b]:k], "densepose": dens})
            return out
        counts, offs, host = ticket["counts"].tolist(), ticket["offs"].tolist(), ticket["host"]
        keys = ("coarse_segm", "fine_segm", "u", "v") + self.big_dev_names[4:]
        for b in range(sess.batch):
            d, o = counts[b], offs[b]
            res = {"image_size": torch.tensor([sess.h0, sess.w0], dtype=torch.int64),
                   "pred_boxes": boxes[b, :d], "scores": scores[b, :d], "pred_classes": torch.zeros(d, dtype=torch.int64)}
            for key, hbuf in zip(keys, host):
                res["pred_densepose_" + key] = hbuf[o:o + d]
            out.append(res)
        return out

    def submit(self, images: torch.Tensor, bgr: bool = True):
        """Enqueue one host batch; returns the oldest finished batch (None while the pipeline fills: the first
        `depth + 1` calls). A pinned `images` tensor is copied to the device directly (the caller keeps it unchanged
        until the results come back); pageable memory goes through the slot's pinned staging buffer first."""
        sl = self.slots[self._next]
        self._next = (self._next + 1) % self.depth
        ticket = self._issue(sl)
        src = images
        if not images.is_pinned():
            sl["host_in"].copy_(images)
            src = sl["host_in"]
        sess = sl["sess"]
        with torch.cuda.device(self.engine.device), torch.cuda.stream(sess.stream):
            sl["dev_in"].copy_(src, non_blocking=True)
            sess.run(sl["dev_in"], bgr)
            for hbuf, dbuf in zip(sl["small_host"], sl["small_dev"]):
                hbuf.copy_(dbuf, non_blocking=True)
            sl["done"].record(sess.stream)
        sl["busy"] = True
        prev = self._finish(self._pending)
        self._pending = ticket
        return prev

    def drain(self) -> List[List[Dict[str, torch.Tensor]]]:
        """Returns every batch still in flight, oldest first (all valid together: depth + 1 result sets)."""
        out = []
        r = self._finish(self._pending)
        self._pending = None
        if r is not None:
            out.append(r)
        for i in range(self.depth):
            sl = self.slots[(self._next + i) % self.depth]
            r = self._finish(self._issue(sl))
            if r is not None:
                out.append(r)
        return out
