"""Torch-tensor front end of the C-ABI kernels (device memory / streams are torch's; the
arithmetic is ours).  Every function here launches hand-written sm_100a kernels through
``_cabi`` — there is no eager / CPU fallback.
"""
import ctypes
from ctypes import c_void_p

import torch

from . import _cabi

_launches = 0          # number of OUR kernel-launching C-ABI calls (bench.py reports it)


def _count(n=1):
    global _launches
    _launches += n


def launch_count() -> int:
    return _launches


class KernelTimer:
    """CUDA-event timing of the masked-GEMM C-ABI calls on the launching stream (bench.py uses it
    to report the dominant kernel's achieved TFLOP/s live, inside the timed region)."""

    def __init__(self):
        self.records = []          # (kind, flops, start_event, end_event)

    def totals(self):
        out = {}
        for kind, flops, e0, e1 in self.records:
            ms = e0.elapsed_time(e1)
            t = out.setdefault(kind, [0.0, 0.0, 0])
            t[0] += ms; t[1] += flops; t[2] += 1
        return out


_timer = None


def set_timer(timer):
    global _timer
    _timer = timer


class _Timed:
    def __init__(self, kind, desc, cin_real=None):
        self.on = _timer is not None
        if self.on:
            cin = desc.cin if cin_real is None else cin_real
            self.flops = 2.0 * desc.n * desc.p * desc.q * desc.cout * cin * desc.r * desc.s
            self.kind = kind

    def __enter__(self):
        if self.on:
            self.e0 = torch.cuda.Event(enable_timing=True); self.e0.record()

    def __exit__(self, *a):
        if self.on:
            e1 = torch.cuda.Event(enable_timing=True); e1.record()
            _timer.records.append((self.kind, self.flops, self.e0, e1))


_grad_ready_hook = None      # set by P2PGradReducer.arm(): called with the data_ptr of every gradient slot a kernel just filled


def set_grad_ready_hook(fn):
    global _grad_ready_hook
    _grad_ready_hook = fn


def get_grad_ready_hook():
    return _grad_ready_hook


def grad_ready(*slots):
    """Report gradients written straight into their persistent slots (no AccumulateGrad node runs for them, so
    autograd's own hooks never fire): lets the gradient exchange start a bucket while the backward pass continues."""
    if _grad_ready_hook is not None:
        for s in slots:
            if s is not None:
                _grad_ready_hook(s.data_ptr())


# ---- weight gradients on a side stream -------------------------------------------------------------------------------
# dW of a layer is needed by nobody before the gradient exchange / the optimizer, while dX is the critical path of the
# backward pass.  With WGRAD_SIDE_STREAM the wgrad GEMM + finalize of every masked layer go to one side stream behind an
# event of the compute stream; at small per-GPU batches (the 8-GPU operating point: 64 images) the kernels are one or two
# waves each and the two chains fill each other's gaps.  The operands are kept alive until ``join_wgrad`` (the caching
# allocator would otherwise hand their blocks to the next main-stream allocation while the side stream still reads them).
# Off unless a caller that also joins turns it on (BaseHarness._step_body does, around its backward pass).
WGRAD_SIDE_STREAM = False
_wgrad_streams = {}
_wgrad_keepalive = []


def set_wgrad_side_stream(on: bool):
    global WGRAD_SIDE_STREAM
    WGRAD_SIDE_STREAM = bool(on)


def _wgrad_stream(device):
    st = _wgrad_streams.get(device.index)
    if st is None:
        st = _wgrad_streams[device.index] = torch.cuda.Stream(device)
    return st


def join_wgrad(device=None):
    """Make the current stream wait for every weight gradient launched on the side stream since the last join, and let go
    of the operands kept alive for them.  Call after ``loss.backward()``, before anything consumes ``param.grad``."""
    if not _wgrad_keepalive:
        return
    devs = {t[0].device for t in _wgrad_keepalive}
    for d in devs:
        if device is None or d == device:
            torch.cuda.current_stream(d).wait_stream(_wgrad_stream(d))
    _wgrad_keepalive.clear()


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("turboprune_b200 kernels need CUDA tensors (B200 / sm_100a); there is no CPU path")


_ws_cache = {}


def _workspace(nbytes: int, device, tag="default") -> torch.Tensor:
    """Grow-only per-(device, stream, tag) scratch buffer owned by Python."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream, tag)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 1 << 16), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


# ---------------------------------------------------------------------------------------------
# pruning
# ---------------------------------------------------------------------------------------------
def topk_threshold_mask(ws, ms, k, gs=None, kind=_cabi.TP_SCORE_MAG, write_masks=True):
    """Exact k-th smallest score over all layers + new masks (torch.kthvalue / torch.where parity).

    Returns (new_masks | None, thr (0-dim fp32 cuda tensor), info dict).  Raises RuntimeError for
    k outside [1, N] like torch.kthvalue does (the reference hits this for k == 0,
    utils/pruning_utils.py:78-79).
    """
    lib = _cabi.load()
    _require_cuda(*ws, *ms)
    dev = ws[0].device
    ws = [w.detach().contiguous() for w in ws]
    ms = [m.detach().contiguous() for m in ms]
    gs = None if gs is None else [g.detach().contiguous() for g in gs]
    for t in ws + ms + (gs or []):
        if t.dtype != torch.float32:
            raise TypeError("pruning kernels operate on fp32 weights / masks / grads")
    numel = [w.numel() for w in ws]
    total = sum(numel)
    outs = [torch.empty_like(m) for m in ms] if write_masks else None
    thr = torch.empty((), dtype=torch.float32, device=dev)
    nbytes = lib.tp_topk_workspace_bytes(len(ws), total)
    wsb = _workspace(nbytes, dev, "topk")
    info = (ctypes.c_int64 * 4)()
    with torch.cuda.device(dev):
        rc = lib.tp_topk_threshold_mask(
            _cabi.ptr_array(ws), _cabi.ptr_array(gs), _cabi.ptr_array(ms), _cabi.ptr_array(outs),
            _cabi.i64_array(numel), len(ws), int(k), int(kind), c_void_p(thr.data_ptr()),
            c_void_p(wsb.data_ptr()), wsb.numel(), info, _cabi.stream_ptr(dev))
    if rc == _cabi.TP_ERR_K_RANGE:
        raise RuntimeError(f"kthvalue(): selected number k out of range for dimension 0 (k={k}, N={total})")
    _cabi.check(rc, "tp_topk_threshold_mask")
    _count(1)
    return outs, thr, {"path": int(info[0]), "candidates": int(info[1]), "n_lt": int(info[2]), "nan_thr": bool(info[3])}


class TopKPlan:
    """Pre-marshalled ``tp_topk_threshold_mask`` call (pointer tables, outputs, workspace built once): ``run(k)`` is
    just the C-ABI call.  Used when the same tensors are pruned repeatedly (benchmarks, per-level IMP)."""

    def __init__(self, ws, ms, gs=None, kind=_cabi.TP_SCORE_MAG):
        self.lib = _cabi.load()
        self.dev = ws[0].device
        self.ws = [w.detach().contiguous() for w in ws]
        self.ms = [m.detach().contiguous() for m in ms]
        self.gs = None if gs is None else [g.detach().contiguous() for g in gs]
        self.outs = [torch.empty_like(m) for m in self.ms]
        self.thr = torch.empty((), dtype=torch.float32, device=self.dev)
        self.n = len(self.ws)
        self.total = sum(w.numel() for w in self.ws)
        self.kind = int(kind)
        self.wsb = torch.empty(self.lib.tp_topk_workspace_bytes(self.n, self.total), dtype=torch.uint8, device=self.dev)
        self.args = (_cabi.ptr_array(self.ws), _cabi.ptr_array(self.gs), _cabi.ptr_array(self.ms), _cabi.ptr_array(self.outs),
                     _cabi.i64_array([w.numel() for w in self.ws]))
        self.info = (ctypes.c_int64 * 4)()
        self._cached = False

    def enqueue(self, k):
        """Issue the whole fast path (one memset + one cooperative kernel) without blocking the stream.  The masks /
        threshold must not be consumed before ``finish(k)``."""
        with torch.cuda.device(self.dev):
            rc = self.lib.tp_topk_enqueue(*self.args, self.n, int(k), self.kind, c_void_p(self.thr.data_ptr()),
                                          c_void_p(self.wsb.data_ptr()), self.wsb.numel(), int(self._cached),
                                          _cabi.stream_ptr(self.dev))
        if rc == _cabi.TP_ERR_K_RANGE:
            raise RuntimeError(f"kthvalue(): selected number k out of range for dimension 0 (k={k}, N={self.total})")
        _cabi.check(rc, "tp_topk_enqueue")
        self._cached = True                      # the segment table now lives in the workspace
        _count(1)

    def finish(self, k):
        """Synchronise, read the status back, run the exact fallback if the bracket missed.  Returns (masks, thr, info)."""
        with torch.cuda.device(self.dev):
            rc = self.lib.tp_topk_finish(self.args[2], self.args[3], self.args[4], self.n, int(k), self.kind,
                                         c_void_p(self.thr.data_ptr()), c_void_p(self.wsb.data_ptr()), self.wsb.numel(),
                                         self.info, _cabi.stream_ptr(self.dev))
        _cabi.check(rc, "tp_topk_finish")
        return self.outs, self.thr, {"path": int(self.info[0]), "candidates": int(self.info[1]), "n_lt": int(self.info[2]),
                                     "nan_thr": bool(self.info[3])}

    def run(self, k):
        self.enqueue(k)
        return self.finish(k)


def apply_threshold(ws, ms, thr, gs=None, kind=_cabi.TP_SCORE_MAG):
    lib = _cabi.load()
    _require_cuda(*ws, *ms, thr)
    dev = ws[0].device
    ws = [w.detach().contiguous() for w in ws]
    ms = [m.detach().contiguous() for m in ms]
    gs = None if gs is None else [g.detach().contiguous() for g in gs]
    outs = [torch.empty_like(m) for m in ms]
    thr = thr.to(device=dev, dtype=torch.float32).reshape(())
    wsb = _workspace(lib.tp_segtable_workspace_bytes(len(ws)), dev, "seg")
    with torch.cuda.device(dev):
        rc = lib.tp_apply_threshold(_cabi.ptr_array(ws), _cabi.ptr_array(gs), _cabi.ptr_array(ms), _cabi.ptr_array(outs),
                                    _cabi.i64_array([w.numel() for w in ws]), len(ws), int(kind),
                                    c_void_p(thr.data_ptr()), c_void_p(wsb.data_ptr()), wsb.numel(), _cabi.stream_ptr(dev))
    _cabi.check(rc, "tp_apply_threshold")
    _count()
    return outs


def count_zeros(ms):
    """int64 cuda tensor [n+1]: zeros per mask and the total in the last slot. One launch, no sync."""
    lib = _cabi.load()
    _require_cuda(*ms)
    dev = ms[0].device
    ms = [m.detach().contiguous() for m in ms]
    out = torch.empty(len(ms) + 1, dtype=torch.int64, device=dev)
    wsb = _workspace(lib.tp_segtable_workspace_bytes(len(ms)), dev, "seg")
    with torch.cuda.device(dev):
        rc = lib.tp_count_zeros(_cabi.ptr_array(ms), _cabi.i64_array([m.numel() for m in ms]), len(ms),
                                c_void_p(out.data_ptr()), c_void_p(wsb.data_ptr()), wsb.numel(), _cabi.stream_ptr(dev))
    _cabi.check(rc, "tp_count_zeros")
    _count()
    return out


# ---------------------------------------------------------------------------------------------
# optimizer
# ---------------------------------------------------------------------------------------------
def sgd_momentum_step(params, grads, bufs, lr_dev, momentum, weight_decay, first_step, table_ws=None, table_cached=False):
    """One fused launch.  ``table_ws``: a persistent uint8 workspace owned by the optimizer; with
    ``table_cached`` the segment table already in it is reused (no H2D copy -> CUDA-graph capturable)."""
    lib = _cabi.load()
    dev = params[0].device
    wsb = table_ws if table_ws is not None else _workspace(lib.tp_segtable_workspace_bytes(len(params)), dev, "seg")
    with torch.cuda.device(dev):
        rc = lib.tp_sgd_momentum(_cabi.ptr_array(params), _cabi.ptr_array(grads), _cabi.ptr_array(bufs),
                                 _cabi.i64_array([p.numel() for p in params]), len(params),
                                 c_void_p(lr_dev.data_ptr()), float(momentum), float(weight_decay), int(bool(first_step)),
                                 int(bool(table_cached)), c_void_p(wsb.data_ptr()), wsb.numel(), _cabi.stream_ptr(dev))
    _cabi.check(rc, "tp_sgd_momentum")
    _count()


# ---------------------------------------------------------------------------------------------
# masked convolution / linear
# ---------------------------------------------------------------------------------------------
def _round_up(x, m):
    return (x + m - 1) // m * m


def make_desc(n, h, w, cin, cout, r, s, stride, padding):
    sh, sw = stride
    ph, pw = padding
    p = (h + 2 * ph - r) // sh + 1
    q = (w + 2 * pw - s) // sw + 1
    return _cabi.ConvDesc(n, h, w, cin, cout, r, s, sh, sw, ph, pw, p, q)


def stem_geometry(cin, r, s):
    """(channels per tap, K of the stem GEMM): no channel padding inside a tap (cg = cin); K padded to a multiple of 8
    (the RGB 7x7 stem: 147 -> 152 columns)."""
    cg = cin
    return cg, _round_up(r * s * cg, 8)


KBLOCK_SKIP = True      # K-block occupancy masks: all-zero 64x64 weight blocks are neither loaded nor multiplied


def set_kblock_skip(on: bool):
    """Toggle tile skipping (benchmarks / parity tests compare both walks; results are bit-identical)."""
    global KBLOCK_SKIP
    KBLOCK_SKIP = bool(on)


def kmask_shapes(cout, cin, r, s, wf_ld, cout_p):
    """([row groups, words] of the fprop mask, [row groups, words] of the dgrad mask) — include/turboprune_b200.h."""
    words = lambda cols: ((cols + 63) // 64 + 31) // 32
    return ((cout + 63) // 64, words(wf_ld)), ((cin + 63) // 64, words(r * s * cout_p))


def kblock_occupancy(kmask, columns):
    """(empty, total) 64x64 weight blocks described by an occupancy mask over ``columns`` K columns (host sync)."""
    kb = (columns + 63) // 64
    rows = kmask_rows(kmask, columns)
    return int(kmask[-1].item()), rows.shape[0] * kb            # the staging call left the count behind the last row


def stage_weights(weight4d, mask4d, cin_p, need_dgrad, cout_p=None, wf_ld=0, want_kmask=None):
    """(mask*w) -> bf16 operand layouts wf [Cout, R*S*cin_p] (row stride ``wf_ld`` when given, zero tail) and
    (optionally) wd [cin, R*S*cout_p].  With ``want_kmask`` (default: the module switch) the K-block occupancy masks are
    produced too and ride along as ``wf.kmask`` / ``wd.kmask`` (uint32 tensors consumed by conv_fprop / conv_dgrad)."""
    lib = _cabi.load()
    cout, cin, r, s = weight4d.shape
    dev = weight4d.device
    if wf_ld and wf_ld > r * s * cin_p:
        wf = torch.zeros(cout, wf_ld, dtype=torch.bfloat16, device=dev)
    else:
        wf = torch.empty(cout, r * s * cin_p, dtype=torch.bfloat16, device=dev)
    wd = None
    cout_p = cout if cout_p is None else cout_p
    if need_dgrad:
        wd = torch.empty(cin, r * s * cout_p, dtype=torch.bfloat16, device=dev)
    kf = kd = None
    if KBLOCK_SKIP if want_kmask is None else want_kmask:
        sf, sd = kmask_shapes(cout, cin, r, s, wf.shape[1], cout_p)
        kf = torch.empty(sf[0] * sf[1] + 1, dtype=torch.int32, device=dev)      # [rows][words] + the empty-block count
        kd = torch.empty(sd[0] * sd[1] + 1, dtype=torch.int32, device=dev) if wd is not None else None
    with torch.cuda.device(dev):
        rc = lib.tp_stage_weights(c_void_p(weight4d.data_ptr()), c_void_p(mask4d.data_ptr()), cout, cin, r, s,
                                  c_void_p(wf.data_ptr()), cin_p, int(wf_ld), c_void_p(wd.data_ptr()) if wd is not None else None,
                                  cout_p, cin, c_void_p(kf.data_ptr()) if kf is not None else None,
                                  c_void_p(kd.data_ptr()) if kd is not None else None, _cabi.stream_ptr(dev))
    _cabi.check(rc, "tp_stage_weights")
    _count()
    wf.kmask = kf
    if wd is not None:
        wd.kmask = kd
    return wf, wd


def kmask_rows(kmask, columns):
    """[row groups, words] view of an occupancy mask buffer (its last element is the empty-block count)."""
    words = ((columns + 63) // 64 + 31) // 32
    return kmask[:-1].view(-1, words)


def padded_cin(cin, r, s):
    """Channel count the TMA layouts need: a multiple of 8 (16-byte rows), and of 64 when the filter has more than one
    tap (the K loop walks 64-channel blocks per tap).  Inputs in between (the reference wraps ANY nn.Conv2d,
    custom_models.py:64-107) are zero-padded to it while being laid out as NHWC bf16."""
    return _round_up(cin, 64 if r * s > 1 else 8)


def _operand_plan(cout, cin, r, s):
    """(cin_p, cout_p, has_wd, wf_ld) of the bf16 operand layouts a masked layer consumes in a train step (the
    WeightStager's view): the stem layout for <= 8 input channels (explicit im2col, no input gradient), otherwise the
    TMA layouts with the input channels padded to ``padded_cin``."""
    if cin <= 8 and (cin % 8 != 0 or r * s > 1):
        cg, kp = stem_geometry(cin, r, s)                    # stem: explicit im2col over padded channel groups, no dgrad
        return cg, cout, False, kp
    cin_p = padded_cin(cin, r, s)
    return cin_p, _round_up(cout, 64 if r * s > 1 else 8), True, r * s * cin_p


class WeightStager:
    """bf16 "weight shadow" of a whole model, refreshed by ONE launch per optimizer step.

    ``stage()`` writes bf16(mask * w) of every masked layer into persistent fprop / dgrad operand buffers and hands
    them to the layers; each layer consumes its pair in its next forward (exactly once) instead of launching its own
    staging kernel — the reference's per-layer ``mask * weight`` + autocast cast (mask_layers.py:25-34) become one
    kernel per step.  Call it right before the training forward; a forward without a preceding ``stage()`` (eval,
    pruning scores) stages per layer as before.  Pointers are re-checked every call (pruning replaces mask tensors),
    the device table is only re-uploaded when they changed, so the call is CUDA-graph capturable after a warm-up step."""

    def __init__(self, layers):
        self.layers = [l for l in layers]
        self._key = None
        self._bufs = None
        self._items = None
        self._ws = None

    @staticmethod
    def _shape4(layer):
        w = layer.weight
        if w.dim() == 4:
            return tuple(w.shape)
        return (w.shape[0], w.shape[1], 1, 1)                 # Linear [out, in] / Conv1d(k=1) [out, in, 1]

    def _rebuild(self, key):
        dev = self.layers[0].weight.device
        if self._bufs is None:
            self._bufs = []
            for l in self.layers:
                cout, cin, r, s = self._shape4(l)
                plan = _operand_plan(cout, cin, r, s)
                if plan is None or not l.weight.is_cuda or l.weight.dtype != torch.float32 or not l.weight.is_contiguous():
                    self._bufs.append(None)
                    continue
                cin_p, cout_p, has_wd, wf_ld = plan
                wf = torch.zeros(cout, wf_ld, dtype=torch.bfloat16, device=dev)
                wd = torch.zeros(cin, r * s * cout_p, dtype=torch.bfloat16, device=dev) if has_wd else None
                self._bufs.append((wf, wd, cin_p, cout_p))
            # K-block occupancy masks of all layers in ONE buffer (zeroed by a single memset node per step)
            shapes = []
            for l, b in zip(self.layers, self._bufs):
                if b is None:
                    shapes.append(None); continue
                cout, cin, r, s = self._shape4(l)
                shapes.append(kmask_shapes(cout, cin, r, s, b[0].shape[1], b[3]))
            total = sum(sf[0] * sf[1] + 1 + ((sd[0] * sd[1] + 1) if b[1] is not None else 0)
                        for (sh, b) in zip(shapes, self._bufs) if sh is not None for sf, sd in [sh])
            self._kmask_all = torch.zeros(max(total, 1), dtype=torch.int32, device=dev)
            off = 0
            for sh, b in zip(shapes, self._bufs):
                if sh is None:
                    continue
                sf, sd = sh
                b[0].kmask = self._kmask_all[off:off + sf[0] * sf[1] + 1]; off += sf[0] * sf[1] + 1
                if b[1] is not None:
                    b[1].kmask = self._kmask_all[off:off + sd[0] * sd[1] + 1]; off += sd[0] * sd[1] + 1
        live = [(l, b) for l, b in zip(self.layers, self._bufs) if b is not None]
        items = (_cabi.StageItem * len(live))()
        for it, (l, (wf, wd, cin_p, cout_p)) in zip(items, live):
            cout, cin, r, s = self._shape4(l)
            it.w = l.weight.data_ptr(); it.mask = l.mask.data_ptr()
            it.wf = wf.data_ptr(); it.wd = wd.data_ptr() if wd is not None else None
            it.cout, it.cin, it.r, it.s, it.cin_p, it.cout_p, it.wf_ld = cout, cin, r, s, cin_p, cout_p, wf.shape[1]
            it.kmask_f = wf.kmask.data_ptr()
            it.kmask_d = wd.kmask.data_ptr() if wd is not None else None
        self._items, self._live, self._key = items, live, key
        if self._ws is None:
            lib = _cabi.load()
            self._ws = torch.empty(max(int(lib.tp_stage_batched_workspace_bytes(len(self.layers))), 256), dtype=torch.uint8, device=dev)

    def stage(self):
        lib = _cabi.load()
        for l in self.layers:                                  # masks created on the host move with the first use
            if l.mask.device != l.weight.device or l.mask.dtype != torch.float32 or not l.mask.is_contiguous():
                l.mask = l.mask.to(device=l.weight.device, dtype=torch.float32).contiguous()
        key = tuple((l.weight.data_ptr(), l.mask.data_ptr()) for l in self.layers)
        cached = key == self._key
        if not cached:
            self._rebuild(key)
        if not len(self._live):
            return
        dev = self.layers[0].weight.device
        with torch.cuda.device(dev):
            rc = lib.tp_stage_weights_batched(self._items, len(self._live), int(cached), c_void_p(self._kmask_all.data_ptr()),
                                              self._kmask_all.numel() * 4, c_void_p(self._ws.data_ptr()),
                                              self._ws.numel(), _cabi.stream_ptr(dev))
        _cabi.check(rc, "tp_stage_weights_batched")
        _count()
        for l, (wf, wd, _, _) in self._live:
            l.__dict__["_tp_staged"] = (wf, wd)


def skipped_block_report(stager):
    """After ``stager.stage()``: per layer and in total, how many 64x64 blocks of the fprop weight operand are empty
    (never loaded / multiplied).  For iid unstructured masks this is ~0 at any density a 64x64 block survives
    (SURVEY.md Appendix B); dead filters / dead input channels are what produces skippable blocks."""
    rows, empty, total = [], 0, 0
    for l, b in zip(stager.layers, stager._bufs):
        if b is None or getattr(b[0], "kmask", None) is None:
            continue
        e, t = kblock_occupancy(b[0].kmask, b[0].shape[1])
        rows.append((type(l).__name__, tuple(l.weight.shape), e, t))
        empty += e; total += t
    return {"empty_blocks": empty, "total_blocks": total, "fraction": empty / max(total, 1), "layers": rows}


def take_staged(layer):
    """The (wf, wd) pair a ``WeightStager`` left for this layer's next forward, or None; consumed exactly once."""
    st = layer.__dict__.get("_tp_staged")
    if st is not None:
        layer.__dict__["_tp_staged"] = None
    return st


def to_nhwc_bf16(x, c_pad):
    """[N, C, H, W] (any strides; fp32 or bf16) -> contiguous NHWC bf16 [N, H, W, c_pad]."""
    lib = _cabi.load()
    n, c, h, w = x.shape
    if x.dtype == torch.bfloat16 and c == c_pad and x.permute(0, 2, 3, 1).is_contiguous():
        return x.permute(0, 2, 3, 1)
    if x.dtype not in (torch.float32, torch.bfloat16):
        x = x.float()
    out = torch.empty(n, h, w, c_pad, dtype=torch.bfloat16, device=x.device)
    sn, sc, sh, sw = x.stride()
    with torch.cuda.device(x.device):
        rc = lib.tp_to_nhwc_bf16(c_void_p(x.data_ptr()), 0 if x.dtype == torch.float32 else 1, sn, sc, sh, sw,
                                 n, c, h, w, c_void_p(out.data_ptr()), c_pad, _cabi.stream_ptr(x.device))
    _cabi.check(rc, "tp_to_nhwc_bf16")
    _count()
    return out


def im2col_c8(x_nhwc8, desc, kp):
    lib = _cabi.load()
    n, h, w, _ = x_nhwc8.shape
    out = torch.empty(n * desc.p * desc.q, kp, dtype=torch.bfloat16, device=x_nhwc8.device)
    with torch.cuda.device(x_nhwc8.device):
        rc = lib.tp_im2col_c8(c_void_p(x_nhwc8.data_ptr()), n, h, w, desc.r, desc.s, desc.stride_h, desc.stride_w,
                              desc.pad_h, desc.pad_w, desc.p, desc.q, c_void_p(out.data_ptr()), kp,
                              _cabi.stream_ptr(x_nhwc8.device))
    _cabi.check(rc, "tp_im2col_c8")
    _count()
    return out


def empty_cl(n, c, h, w, device):
    """[n, c, h, w] bf16 with channels_last strides: the memory IS an NHWC array, and the tensor is not a view
    (autograd forbids in-place ops such as nn.ReLU(inplace=True) on views created inside a custom Function)."""
    return torch.empty((n, c, h, w), dtype=torch.bfloat16, device=device, memory_format=torch.channels_last)


def im2col_stem(x, desc, kp, cg):
    """[N, C<=8, H, W] fp32/bf16 (any strides) -> [N*P*Q, kp] bf16 im2col matrix (column tap*cg + channel), conversion fused in."""
    lib = _cabi.load()
    n, c, h, w = x.shape
    if x.dtype not in (torch.float32, torch.bfloat16):
        x = x.float()
    out = torch.empty(n * desc.p * desc.q, kp, dtype=torch.bfloat16, device=x.device)
    sn, sc, sh, sw = x.stride()
    with torch.cuda.device(x.device):
        rc = lib.tp_im2col_stem(c_void_p(x.data_ptr()), 0 if x.dtype == torch.float32 else 1, sn, sc, sh, sw, n, c, h, w,
                                desc.r, desc.s, cg, desc.stride_h, desc.stride_w, desc.pad_h, desc.pad_w, desc.p, desc.q,
                                c_void_p(out.data_ptr()), kp, _cabi.stream_ptr(x.device))
    _cabi.check(rc, "tp_im2col_stem")
    _count()
    return out


def conv_fprop(desc, x_nhwc, wf, bias=None, out=None, want_stats=False):
    """``want_stats``: also return the BatchNorm batch statistics of the output, written by the conv epilogue
    ([rows, 2, cout] fp32: per 32-pixel group and channel the sum and the sum of squares of the bf16 outputs)."""
    lib = _cabi.load()
    dev = x_nhwc.device
    y = out if out is not None else torch.empty(desc.n, desc.p, desc.q, desc.cout, dtype=torch.bfloat16, device=dev)
    stats = None
    if want_stats:
        stats = torch.empty(int(lib.tp_conv_stats_rows(ctypes.byref(desc))), 2, desc.cout, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev), _Timed("fprop", desc):
        km = getattr(wf, "kmask", None) if KBLOCK_SKIP else None
        rc = lib.tp_conv_fprop_stats(ctypes.byref(desc), c_void_p(x_nhwc.data_ptr()), c_void_p(wf.data_ptr()),
                                     c_void_p(km.data_ptr()) if km is not None else None,
                                     c_void_p(bias.data_ptr()) if bias is not None else None, c_void_p(y.data_ptr()),
                                     c_void_p(stats.data_ptr()) if stats is not None else None,
                                     None, 0, _cabi.stream_ptr(dev))
    _cabi.check(rc, "tp_conv_fprop")
    _count()
    return (y, stats) if want_stats else y


def conv_dgrad(desc, dy_nhwc, wd, addend=None, kmask=None):
    """dx = dgrad(dy) [+ addend]: ``addend`` (NHWC bf16, dx's shape) is accumulated in the kernel epilogue.
    ``kmask``: K-block occupancy mask of ``wd`` (defaults to the one riding on the tensor)."""
    lib = _cabi.load()
    dev = dy_nhwc.device
    dx = torch.empty(desc.n, desc.h, desc.w, desc.cin, dtype=torch.bfloat16, device=dev)
    with torch.cuda.device(dev), _Timed("dgrad", desc):
        km = (kmask if kmask is not None else getattr(wd, "kmask", None)) if KBLOCK_SKIP else None
        rc = lib.tp_conv_dgrad(ctypes.byref(desc), c_void_p(dy_nhwc.data_ptr()), c_void_p(wd.data_ptr()),
                               c_void_p(km.data_ptr()) if km is not None else None,
                               c_void_p(addend.data_ptr()) if addend is not None else None,
                               c_void_p(dx.data_ptr()), None, 0, _cabi.stream_ptr(dev))
    _cabi.check(rc, "tp_conv_dgrad")
    _count(1 if desc.stride_h == 1 and desc.stride_w == 1 else desc.stride_h * desc.stride_w)
    return dx


BN_BWD_FUSION = True     # dgrad epilogue does the BatchNorm backward reduction of the layer that feeds the convolution


def set_bn_bwd_fusion(on: bool):
    global BN_BWD_FUSION
    BN_BWD_FUSION = bool(on)


def conv_dgrad_bnrelu(desc, dy_nhwc, wd, bn_src, kmask=None):
    """dgrad whose result is the gradient of a BatchNorm+ReLU output: returns (g, partial) with g = dx * [z > 0] and the
    per-32-pixel partial sums (sum g, sum g * xhat) the BatchNorm backward needs; None when the shape has no staged path."""
    lib = _cabi.load()
    dev = dy_nhwc.device
    y, weight, bias, mean, invstd = bn_src[:5]
    g = torch.empty(desc.n, desc.h, desc.w, desc.cin, dtype=torch.bfloat16, device=dev)
    rows = int(lib.tp_conv_dgrad_partial_rows(ctypes.byref(desc)))
    partial = torch.empty(rows, 2, desc.cin, dtype=torch.float32, device=dev)
    km = (kmask if kmask is not None else getattr(wd, "kmask", None)) if KBLOCK_SKIP else None
    P = lambda t: c_void_p(t.data_ptr()) if t is not None else None
    with torch.cuda.device(dev), _Timed("dgrad", desc):
        rc = lib.tp_conv_dgrad_bnrelu(ctypes.byref(desc), P(dy_nhwc), P(wd), P(km), P(y), P(weight), P(bias), P(mean), P(invstd),
                                      P(g), P(partial), _cabi.stream_ptr(dev))
    if rc == -5:          # TP_ERR_UNSUPPORTED: no 16-byte aligned linear output for this shape
        return None
    _cabi.check(rc, "tp_conv_dgrad_bnrelu")
    _count()
    return g, partial


def conv_wgrad(desc, x_nhwc, dy_nhwc, mask4d, cin_real, want_db=False, dw_out=None, db_out=None, kmask=None):
    """``dw_out`` / ``db_out``: write the gradients straight into these (contiguous fp32) buffers — used with the
    persistent gradient arena so no separate accumulate kernel runs.  ``kmask``: the fprop occupancy mask staged for
    ``mask4d`` (``wf.kmask``): output tiles under all-zero mask blocks are skipped (their gradient is zero)."""
    lib = _cabi.load()
    dev = x_nhwc.device
    dw = dw_out if dw_out is not None else torch.empty(desc.cout, cin_real, desc.r, desc.s, dtype=torch.float32, device=dev)
    db = (db_out if db_out is not None else torch.empty(desc.cout, dtype=torch.float32, device=dev)) if want_db else None
    nbytes = lib.tp_conv_workspace_bytes(ctypes.byref(desc), 2)
    wsb = _workspace(nbytes, dev, "wgrad")
    with torch.cuda.device(dev), _Timed("wgrad", desc):
        km = kmask if KBLOCK_SKIP else None
        rc = lib.tp_conv_wgrad(ctypes.byref(desc), c_void_p(x_nhwc.data_ptr()), c_void_p(dy_nhwc.data_ptr()),
                               c_void_p(mask4d.data_ptr()), c_void_p(km.data_ptr()) if km is not None else None,
                               cin_real, c_void_p(dw.data_ptr()),
                               c_void_p(db.data_ptr()) if db is not None else None,
                               c_void_p(wsb.data_ptr()), wsb.numel(), _cabi.stream_ptr(dev))
    _cabi.check(rc, "tp_conv_wgrad")
    _count(3 if want_db else 2)
    return dw, db


class MaskedConv2dFn(torch.autograd.Function):
    """y = conv2d(x, mask*w, b) with bf16 tensor-core operands and fp32 accumulation.

    Replaces ``F.conv2d(x, mask * weight, ...)`` under bf16 autocast (reference
    utils/mask_layers.py:23-34 executed inside base_harness.py:121-125) and its autograd
    backward: dX = dgrad(dY, mask*w), dW = mask * wgrad(x, dY) in fp32, db = sum dY.
    Activations stay NHWC bf16 (logical NCHW tensors with channels_last strides).
    """

    @staticmethod
    def forward(ctx, x, weight, mask, bias, stride, padding, want_skip=False, grad_slots=None, staged=None, want_stats=False,
                bn_src=None):
        _require_cuda(x, weight, mask)
        ctx.set_materialize_grads(False)
        ctx.bn_src = None
        ctx.want_skip = want_skip
        ctx.want_stats = want_stats
        stats = None
        # (w_slot, b_slot): persistent arena slots; when given, backward writes dW / db there and returns None for
        # them (no AccumulateGrad add kernel; the slot IS param.grad)
        ctx.grad_slots = grad_slots
        cout, cin, r, s = weight.shape
        n, _, h, w = x.shape
        need_dx = ctx.needs_input_grad[0]
        w32 = weight.detach().contiguous()
        m32 = mask.detach().contiguous()
        cin_p = padded_cin(cin, r, s)
        # <= 8 input channels and no input gradient wanted (the RGB stem): explicit im2col + plain GEMM, K = taps * cin.
        # Everything else goes through the TMA layouts with the channels zero-padded to cin_p.
        small_c = cin <= 8 and cin_p != cin and not need_dx
        desc = make_desc(n, h, w, cin_p if not small_c else cin, cout, r, s, stride, padding)
        # the backward GEMMs contract over Cout: filters larger than 1x1 walk it in 64-channel blocks per tap
        cout_p = _round_up(cout, 64 if (r * s > 1 and not small_c) else 8)    # (the stem path is a plain GEMM over im2col)
        if small_c:
            # stem conv: cg = cin channels per tap, explicit im2col, then a plain GEMM
            cg, kp = stem_geometry(cin, r, s)
            xg = im2col_stem(x, desc, kp, cg)
            gdesc = _cabi.ConvDesc(n * desc.p * desc.q, 1, 1, kp, cout, 1, 1, 1, 1, 0, 0, 1, 1)
            if staged is not None and staged[0].shape == (cout, kp):
                wf, wd = staged
            else:
                wf, wd = stage_weights(w32, m32, cg, False, wf_ld=kp)
            y = empty_cl(n, cout, desc.p, desc.q, x.device)
            if want_stats:
                _, stats = conv_fprop(gdesc, xg, wf, bias, out=y, want_stats=True)
            else:
                conv_fprop(gdesc, xg, wf, bias, out=y)
            ctx.mode = "stem"
            ctx.gdesc = gdesc
            ctx.save_for_backward(xg, m32)
        else:
            xn = to_nhwc_bf16(x, cin_p)      # channels cin..cin_p are zero (and so are the staged weights there)
            if (staged is not None and staged[0].shape == (cout, r * s * cin_p)
                    and (not need_dx or (staged[1] is not None and staged[1].shape == (cin, r * s * cout_p)))):
                wf, wd = staged              # refreshed by WeightStager.stage() for this step (one launch for all layers)
            else:
                wf, wd = stage_weights(w32, m32, cin_p, need_dx, cout_p)
            y = empty_cl(n, cout, desc.p, desc.q, x.device)
            if want_stats:
                _, stats = conv_fprop(desc, xn, wf, bias, out=y, want_stats=True)
            else:
                conv_fprop(desc, xn, wf, bias, out=y)
            ctx.mode = "conv"
            ctx.save_for_backward(xn, m32, wd)
            # x is the output of a fused BatchNorm+ReLU (no residual): this layer's dgrad can do that BatchNorm's backward
            # reduction in its epilogue (stride 1, no channel padding, the NHWC buffers line up)
            if (bn_src is not None and BN_BWD_FUSION and need_dx and not want_skip and stride == (1, 1) and cin_p == cin
                    and bn_src[0].shape == xn.shape):
                ctx.bn_src = bn_src
            ctx.wd_kmask = getattr(wd, "kmask", None) if wd is not None else None     # attributes do not survive save_for_backward
            ctx.wf_kmask = getattr(wf, "kmask", None)                                  # wgrad skips tiles under all-zero mask blocks
        ctx.desc = desc
        ctx.cin = cin
        ctx.has_bias = bias is not None
        ctx.cout_p = cout_p
        ctx.x_dtype = x.dtype
        if want_skip:
            # second output = the input itself: whatever gradient reaches it (the identity path of a residual
            # block, or a downsample branch) comes back to backward() as ``dskip`` and is accumulated inside
            # the dgrad epilogue instead of by autograd's separate elementwise add
            if want_stats:
                ctx.mark_non_differentiable(stats)
                return y, x, stats
            return y, x
        if want_stats:
            ctx.mark_non_differentiable(stats)
            return y, stats
        return y

    @staticmethod
    def backward(ctx, dy, *rest):
        # outputs were (y [, x_skip] [, stats]); the statistics output is non-differentiable
        dskip = rest[0] if ctx.want_skip and rest else None
        desc = ctx.desc
        if dy is None:          # only the skip output was used downstream
            return dskip, None, None, None, None, None, None, None, None, None, None
        cout = desc.cout
        need_dx, need_dw = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        need_db = ctx.has_bias and ctx.needs_input_grad[3]
        dyn = to_nhwc_bf16(dy, ctx.cout_p)
        dx = dw = db = None
        db_in_slot = False          # the bias gradient already sits in param.grad's arena slot: return None for it
        if ctx.mode == "stem":
            xg, m32 = ctx.saved_tensors
            if need_dw:
                r, s = desc.r, desc.s
                cin = m32.shape[1]
                ones = torch.ones(cout, xg.shape[1], dtype=torch.float32, device=xg.device)
                gd = ctx.gdesc
                dwm, db = conv_wgrad(gd, xg, dyn.view(-1, cout), ones.view(cout, -1, 1, 1), xg.shape[1], need_db)
                # columns are (tap, channel group): back to OIHW and apply the mask (9.4 k elements)
                cg = stem_geometry(cin, r, s)[0]
                dw = dwm[:, :r * s * cg].reshape(cout, r * s, cg)[:, :, :cin].permute(0, 2, 1).reshape(cout, cin, r, s) * m32
        else:
            xn, m32, wd = ctx.saved_tensors
            cin = ctx.cin                      # real input channels (desc.cin is the padded count the activation carries)
            ddesc = desc
            if ctx.cout_p != cout:
                ddesc = _cabi.ConvDesc(desc.n, desc.h, desc.w, desc.cin, ctx.cout_p, desc.r, desc.s, desc.stride_h,
                                       desc.stride_w, desc.pad_h, desc.pad_w, desc.p, desc.q)
            if need_dx:
                addend = None
                if dskip is not None:
                    addend = to_nhwc_bf16(dskip, cin)
                # dX has the REAL channel count: the dgrad GEMM's N axis is cin, its K axis (taps x cout_p)
                xdesc = ddesc if cin == desc.cin else _cabi.ConvDesc(desc.n, desc.h, desc.w, cin, ctx.cout_p, desc.r, desc.s,
                                                                      desc.stride_h, desc.stride_w, desc.pad_h, desc.pad_w,
                                                                      desc.p, desc.q)
                fused = None
                if ctx.bn_src is not None and addend is None:
                    fused = conv_dgrad_bnrelu(xdesc, dyn, wd, ctx.bn_src, kmask=ctx.wd_kmask)
                if fused is not None:
                    g, partial = fused
                    from . import fused_norm
                    fused_norm.offer_partials(g, partial, ctx.bn_src[5])          # picked up by that BatchNorm's backward
                    dx = g.permute(0, 3, 1, 2)
                else:
                    dx = conv_dgrad(xdesc, dyn, wd, addend, kmask=ctx.wd_kmask).permute(0, 3, 1, 2)
                if dx.dtype != ctx.x_dtype:
                    dx = dx.to(ctx.x_dtype)
            if need_dw:
                if ctx.cout_p != cout:
                    m_p = torch.zeros(ctx.cout_p, *m32.shape[1:], dtype=torch.float32, device=m32.device)
                    m_p[:cout] = m32
                    dwp, dbp = conv_wgrad(ddesc, xn, dyn, m_p, cin, need_db)
                    dw = dwp[:cout].contiguous()
                    db = dbp[:cout].contiguous() if dbp is not None else None
                else:
                    ws_, bs_ = ctx.grad_slots if ctx.grad_slots is not None else (None, None)
                    direct_w = ws_ is not None and ws_.is_contiguous() and ws_.numel() == m32.numel()
                    direct_b = need_db and bs_ is not None
                    if WGRAD_SIDE_STREAM and direct_w and (direct_b or not need_db):
                        # nothing of this result flows back through autograd: run it beside the dgrad chain
                        dev = xn.device
                        cur, side = torch.cuda.current_stream(dev), _wgrad_stream(dev)
                        ev = torch.cuda.Event(); ev.record(cur)
                        side.wait_event(ev)
                        with torch.cuda.stream(side):
                            conv_wgrad(desc, xn, dyn, m32, cin, need_db, dw_out=ws_, db_out=bs_ if direct_b else None,
                                       kmask=ctx.wf_kmask)
                            grad_ready(ws_, bs_ if direct_b else None)
                        _wgrad_keepalive.append((xn, dyn, m32))
                        dw = db = None
                        db_in_slot = direct_b
                    else:
                        dw, db = conv_wgrad(desc, xn, dyn, m32, cin, need_db, dw_out=ws_ if direct_w else None,
                                            db_out=bs_ if direct_b else None, kmask=ctx.wf_kmask)
                        if direct_w:
                            dw = None
                        if direct_b:
                            db = None
                            db_in_slot = True
                        grad_ready(ws_ if direct_w else None, bs_ if direct_b else None)
        if need_db and db is None and not db_in_slot:
            db = dy.float().sum(dim=(0, 2, 3))
        if dskip is not None and dx is None and need_dx is False:
            dx = None
        return dx, dw, None, db, None, None, None, None, None, None, None


def masked_conv2d(x, weight, mask, bias=None, stride=(1, 1), padding=(0, 0), want_skip=False, grad_slots=None, staged=None,
                  want_stats=False, bn_src=None):
    """Returns y, or (y, x_skip) with ``want_skip``, with the BatchNorm statistics tensor appended for ``want_stats``.
    ``bn_src``: what ``BatchNorm2dB200`` attaches to its BatchNorm+ReLU output (``x._tp_bn_src``)."""
    return MaskedConv2dFn.apply(x, weight, mask, bias, tuple(stride), tuple(padding), want_skip, grad_slots, staged, want_stats,
                                bn_src)


def masked_linear(x, weight2d, mask2d, bias=None, grad_slots=None, staged=None):
    """y = x @ (mask*w)^T + b for x [..., in]; runs as a 1x1 convolution over a [rows,1,1,in] image."""
    shp = x.shape
    x2 = x.reshape(-1, shp[-1])
    y = MaskedConv2dFn.apply(x2.view(x2.shape[0], x2.shape[1], 1, 1), weight2d.view(*weight2d.shape, 1, 1),
                             mask2d.view(*mask2d.shape, 1, 1), bias, (1, 1), (0, 0), False, grad_slots, staged)
    return y.reshape(*shp[:-1], weight2d.shape[0])
