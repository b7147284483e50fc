"""Train-mode execution of the hot path: forward with batch-statistics BatchNorm and the full backward
(reference: ``train.py:153-160`` - ``model.train()``, ``loss.backward()`` through
``retrain/LEAStereo.py:34-51``, ``retrain/skip_model_3d.py``, ``models/operations_3d.py:41-47``,
``models/build_model_2d.py:52-57``).

Correctness-first design: every layer is one node with a ``forward`` and a ``backward`` over pre-allocated planes
volumes (value buffer + gradient buffer per activation).  ConvBR in train mode = conv (raw, tensor-core or SIMT kernel)
-> per-channel batch statistics (``lea_channel_reduce``) -> normalise + ReLU (+ state-sum accumulate)
(``lea_affine_relu``); its backward = two per-channel reductions + ``lea_bn_relu_bwd`` -> data gradient (the forward
conv kernel with transposed, tap-flipped weights, accumulated into the input's gradient) + weight gradient
(``lea_conv3d_wgrad``).  BatchNorm statistics are per replica (what nn.DataParallel gives the reference);
running statistics follow momentum 0.1 / unbiased variance.  There is no autograd/PyTorch fallback for any of this.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch

from .engine import Slice, _prod
from .kernels import LeaError, Ops, PlanesVol, get_ops, lea_tc_opts
from .modules import ConvBR3d, Identity3d, newMatching
from .structure import scale_dimension


class _Node:
    def forward(self):
        raise NotImplementedError

    def backward(self):
        raise NotImplementedError

    def plan_backward(self, cover: "_Coverage"):
        """Declares, in backward order, which gradient slices this node reads and writes (``TrainPlan._finalize``)."""
        raise NotImplementedError


class _Coverage:
    """Which channel ranges of every gradient volume have been written so far while the backward order is walked.
    The first writer of a range overwrites (no zero fill of the volume, no read-modify-write); a volume whose writes
    overlap only partially, or that is read before it is completely written, is zero-filled at the start of the
    backward instead and all its writers accumulate."""

    def __init__(self, zero_vols):
        self.zero_vols = zero_vols                # ids of the volumes that are zero-filled: everything accumulates
        self.ranges: Dict[int, List[Tuple[int, int]]] = {}
        self.retry = False

    def _covered(self, vid, c0, c) -> Tuple[bool, bool]:
        """(any overlap, fully covered) of [c0, c0 + c) against the ranges written so far."""
        rs = sorted(self.ranges.get(vid, []))
        overlap = any(r0 < c0 + c and c0 < r0 + rc for r0, rc in rs)
        pos = c0
        for r0, rc in rs:
            if r0 <= pos < r0 + rc:
                pos = r0 + rc
        return overlap, pos >= c0 + c

    def write(self, sl: Slice) -> bool:
        """Registers a write of the gradient slice; returns True when the writer has to accumulate."""
        vid = id(sl.vol)
        if vid in self.zero_vols:
            return True
        overlap, full = self._covered(vid, sl.c0, sl.c)
        if not overlap:
            self.ranges.setdefault(vid, []).append((sl.c0, sl.c))
            return False
        if full:
            return True
        self.zero_vols.add(vid)
        self.retry = True
        return True

    def read(self, sl: Slice):
        vid = id(sl.vol)
        if vid in self.zero_vols:
            return
        if not self._covered(vid, sl.c0, sl.c)[1]:
            self.zero_vols.add(vid)
            self.retry = True


class TrainPlan:
    def __init__(self, matching: newMatching, ops: Ops, B: int, spatial, planes: int, device, conv_mode: str, maxdisp):
        self.m, self.ops, self.B, self.spatial, self.P = matching, ops, B, tuple(int(v) for v in spatial), planes
        self.device = torch.device(device)
        self.conv_mode = conv_mode
        self.maxdisp = maxdisp
        self.nodes: List[_Node] = []
        self.grads: Dict[int, PlanesVol] = {}
        self.values: List[PlanesVol] = []
        self.param_grads: Dict[torch.nn.Parameter, torch.Tensor] = {}
        self._resampled: Dict[tuple, Slice] = {}
        self.generation = 0                        # forward counter: a backward must belong to the latest forward
        # CUDA graphs of the two launch lists (~1100 + ~1600 launches through ctypes: the eager step is host-bound).
        # Two eager steps first (they create the per-node weight-image buffers), then capture; the graphs are dropped
        # when a parameter moves to other storage (e.g. an optimizer that re-points them into a flat buffer).
        self.use_graph = False
        self._eager_calls = [0, 0]
        self._graphs = [None, None]
        self._graph_key = None
        self._build()
        self._finalize()

    # ---- buffers ----------------------------------------------------------------------------------------
    def vol(self, c, spatial) -> PlanesVol:
        v = PlanesVol.empty(self.B, c, self.P, *spatial, self.device)
        self.values.append(v)
        return v

    def grad_of(self, v: PlanesVol) -> PlanesVol:
        g = self.grads.get(id(v))
        if g is None:
            g = PlanesVol.empty(v.B, v.C, v.P, v.D, v.H, v.W, self.device)
            self.grads[id(v)] = g
        return g

    def _finalize(self):
        """Parameter-gradient bucket (one flat fp32 buffer, zeroed by ONE memset per backward; the weight-gradient and
        BN kernels add into views of it) and the first-writer analysis of the activation gradients."""
        params = used_parameters(self)
        total = sum(p.numel() for p in params)
        self.pflat = torch.zeros(max(1, total), dtype=torch.float32, device=self.device)
        off = 0
        self.param_grads = {}
        for p in params:
            self.param_grads[p] = self.pflat[off:off + p.numel()].view(p.shape)
            off += p.numel()
        for v in self.values:
            self.grad_of(v)
        zero_vols: set = set()
        for _ in range(len(self.values) + 2):
            cover = _Coverage(zero_vols)
            for n in reversed(self.nodes):
                n.plan_backward(cover)
            cover.read(Slice(self.cost, 0, self.cost.C))
            if not cover.retry:
                break
        else:
            raise LeaError("training plan: gradient coverage analysis did not converge")
        self.zero_grads = [self.grads[id(v)] for v in self.values if id(v) in zero_vols]

    # ---- conv launch helper (raw conv: no BN, no ReLU) ----------------------------------------------------
    def conv_raw(self, src: Slice, weight: torch.Tensor, dst: Optional[Slice], accumulate: bool = False,
                 dst_f32: Optional[torch.Tensor] = None, dgrad_of: Optional[torch.Tensor] = None, images=None):
        """``weight`` (c_out, c_in, k, k, k), or - ``dgrad_of`` = a FORWARD weight (c_in, c_out, k, k, k) - the data-gradient
        conv of that layer (channels transposed, taps flipped; packed in place on the tensor-core path).  ``images`` = the
        caller's list of packed-weight buffers, one per 64-channel output chunk (created on first use)."""
        if dgrad_of is not None:
            c_in, c_out, k = dgrad_of.shape[0], dgrad_of.shape[1], dgrad_of.shape[2]
        else:
            c_out, c_in, k = weight.shape[0], weight.shape[1], weight.shape[2]
            weight = weight.contiguous()
        if src.c != c_in:
            raise LeaError("conv_raw: %d input channels, weight expects %d" % (src.c, c_in))
        # the kernels take at most 64 output channels per launch
        for n, o0 in enumerate(range(0, c_out, 64)):
            oc = min(64, c_out - o0)
            d = None if dst is None else Slice(dst.vol, dst.c0 + o0, oc)
            p = self.ops.make_conv(src.vol, src.c0, c_in, oc, k, None, None, False,
                                   dst=None if d is None else d.vol, dst_c0=0 if d is None else d.c0,
                                   res=d.vol if (accumulate and d is not None) else None,
                                   res_c0=d.c0 if (accumulate and d is not None) else 0, dst_f32=dst_f32)
            tc = self.conv_mode == "tc" and self.ops.tc_weight_image_bytes(c_in, oc, k, self.P) > 0
            if images is not None and len(images) <= n:
                images.append(None)
            out = images[n] if images is not None else None
            if tc and dgrad_of is not None:
                img = self.ops.pack_weights_tc_dgrad(dgrad_of, self.P, o0, oc, out=out)
            else:
                if weight is None:               # SIMT path of a data gradient: materialise the transposed weight once
                    weight = dgrad_of.flip(2, 3, 4).transpose(0, 1).contiguous()
                w = weight[o0:o0 + oc].contiguous() if (o0 > 0 or oc < c_out) else weight
                img = self.ops.pack_weights_tc(w, self.P, out=out) if tc else None
            if tc:
                if images is not None:
                    images[n] = img
                self.ops.conv3d_tc(p, img, lea_tc_opts(), src.vol.t)
            else:
                self.ops.conv3d_simt(p, w, src.vol.t)

    # ---- graph construction (unfused mirror of retrain/skip_model_3d.py) -----------------------------------
    def _convbr(self, name, mod: ConvBR3d, src: Slice, dst: Slice, accumulate=False, upsample_to=None):
        self.nodes.append(_ConvBRNode(self, name, mod, src, dst, accumulate, upsample_to))

    def _resample(self, name, src: Slice, spatial) -> Slice:
        """One resample per distinct (tensor, size): s1 of cell i is s0 of cell i+1 (skip_model_3d.py:44-51)."""
        key = (id(src.vol), src.c0, src.c, tuple(spatial))
        dst = self._resampled.get(key)
        if dst is None:
            dst = Slice(self.vol(src.c, spatial), 0, src.c)
            self.nodes.append(_ResampleNode(self, name, src, dst))
            self._resampled[key] = dst
        return dst

    def _pre(self, name, mod: ConvBR3d, src: Slice, spatial, dst: Slice):
        """Cell input path [resample ->] 1x1x1 ConvBR (skip_model_3d.py:44-53).  When the resample ENLARGES the volume
        the conv runs first on the small volume - a 1x1x1 conv commutes with the interpolation - and its raw output is
        up-sampled before the batch statistics: c_out instead of c_in channels at the high resolution (8 instead of 64
        for cell 10), in the forward and in every backward kernel."""
        k = mod.conv.weight.shape[2]
        if src.spatial == tuple(spatial):
            self._convbr(name, mod, src, dst)
        elif k == 1 and _prod(src.spatial) < _prod(spatial):
            self._convbr(name, mod, src, dst, upsample_to=tuple(spatial))
        else:
            self._convbr(name, mod, self._resample(name + ".resample", src, spatial), dst)

    def _cell(self, i, s0: Slice, s1: Slice, out: Optional[Slice] = None):
        cell = self.m.cells[i]
        spec = cell.spec
        name = "cells.%d" % i
        prev_input = s1
        c_out = spec.c_out
        sp = s1.spatial
        if spec.downup_sample != 0:
            sp = tuple(scale_dimension(n, spec.scale) for n in s1.spatial)
        bm = self.m._block_multiplier
        n_states = 2 + len(spec.steps)
        first = n_states - bm
        if out is None:
            out = Slice(self.vol(bm * c_out, sp), 0, bm * c_out)
        elif out.spatial != sp or out.c != bm * c_out:
            raise LeaError("cell %d output does not fit its skip-concat slot" % i)

        def slot(q):
            pos = q - first
            return Slice(out.vol, out.c0 + pos * c_out, c_out) if pos >= 0 else Slice(self.vol(c_out, sp), 0, c_out)

        if s0.c != c_out:
            d0 = slot(0)
            self._pre(name + ".pre_preprocess", cell.pre_preprocess, s0, sp, d0)
            s0 = d0
        else:
            if s0.spatial != sp:
                s0 = self._resample(name + ".resample_s0", s0, sp)
            if first <= 0:
                d0 = slot(0)
                self.nodes.append(_CopyNode(self, name + ".s0_copy", s0, d0, False))
                s0 = d0
        d1 = slot(1)
        self._pre(name + ".preprocess", cell.preprocess, s1, sp, d1)
        states = [s0, d1]
        for k, step in enumerate(spec.steps):
            dst = slot(2 + k)
            for n, (j, opi) in enumerate(step):
                op = cell._ops[opi]
                if isinstance(op, Identity3d):
                    self.nodes.append(_CopyNode(self, "%s._ops.%d(skip)" % (name, opi), states[j], dst, n > 0))
                else:
                    self._convbr("%s._ops.%d" % (name, opi), op, states[j], dst, accumulate=n > 0)
            states.append(dst)
        return prev_input, out

    def _build(self):
        m = self.m
        fm = m.initial_fm
        L0 = self.spatial
        self.cost = self.vol(2 * fm, L0)
        v0, v1 = self.vol(fm, L0), self.vol(fm, L0)
        self._convbr("stem0", m.stem0, Slice(self.cost, 0, 2 * fm), Slice(v0, 0, fm))
        self._convbr("stem1", m.stem1, Slice(v0, 0, fm), Slice(v1, 0, fm))
        stem0, stem1 = Slice(v0, 0, fm), Slice(v1, 0, fm)
        out0 = self._cell(0, stem0, stem1)
        spec1 = m.cells[1].spec
        sp1 = out0[1].spatial if spec1.downup_sample == 0 else \
            tuple(scale_dimension(n, spec1.scale) for n in out0[1].spatial)
        cw = m._block_multiplier * spec1.c_out
        skip = self.vol(3 * cw, sp1)
        out1 = self._cell(1, out0[0], out0[1], Slice(skip, 0, cw))
        out2 = self._cell(2, out1[0], out1[1])
        out3 = self._cell(3, out2[0], out2[1])
        out4 = self._cell(4, out3[0], out3[1], Slice(skip, cw, cw))
        x5 = Slice(self.vol(m.conv1.conv.out_channels, sp1), 0, m.conv1.conv.out_channels)
        self._convbr("conv1", m.conv1, Slice(skip, 0, 2 * cw), x5)
        out5 = self._cell(5, out4[0], x5)
        out6 = self._cell(6, out5[0], out5[1])
        out7 = self._cell(7, out6[0], out6[1])
        out8 = self._cell(8, out7[0], out7[1], Slice(skip, 2 * cw, cw))
        x9 = Slice(self.vol(m.conv2.conv.out_channels, sp1), 0, m.conv2.conv.out_channels)
        self._convbr("conv2", m.conv2, Slice(skip, cw, 2 * cw), x9)
        out9 = self._cell(9, out8[0], x9)
        out10 = self._cell(10, out9[0], out9[1])
        out11 = self._cell(11, out10[0], out10[1])
        last = out11[1]
        d, h, w = L0

        def conv_to(name, mod, src):
            dst = Slice(self.vol(mod.conv.out_channels, src.spatial), 0, mod.conv.out_channels)
            self._convbr(name, mod, src, dst)
            return dst

        if last.spatial[1] == h:
            feat = last
        elif last.spatial[1] == h // 2:
            feat = self._resample("head.upsample_6", conv_to("last_6", m.last_6, last), L0)
        elif last.spatial[1] == h // 4:
            t = self._resample("head.upsample_12", conv_to("last_12", m.last_12, last), (d // 2, h // 2, w // 2))
            feat = self._resample("head.upsample_6", conv_to("last_6", m.last_6, t), L0)
        elif last.spatial[1] == h // 8:
            t = self._resample("head.upsample_24", conv_to("last_24", m.last_24, last), (d // 4, h // 4, w // 4))
            t = self._resample("head.upsample_12", conv_to("last_12", m.last_12, t), (d // 2, h // 2, w // 2))
            feat = self._resample("head.upsample_6", conv_to("last_6", m.last_6, t), L0)
        else:
            raise LeaError("matching net ends on a level the reference head does not handle")
        self.mat = torch.empty((self.B, 1, d, h, w), dtype=torch.float32, device=self.device)
        self.nodes.append(_Last3Node(self, "last_3", m.last_3, feat))

    # ---- execution ----------------------------------------------------------------------------------------
    def forward(self, fx: torch.Tensor, fy: torch.Tensor) -> torch.Tensor:
        self.generation += 1                       # the plan's buffers now hold THIS forward's activations
        self.ops.cost_volume_planes(fx, fy, self.maxdisp, self.P, out=self.cost)
        for n in self.nodes:
            n.forward()
        return self.ops.disp_head(self.mat, self.maxdisp)

    def _storage_key(self):
        return tuple(p.data_ptr() for p in self.param_grads) + tuple(
            t.data_ptr() for n in self.nodes if isinstance(n, _ConvBRNode) and n.mod.bn.running_mean is not None
            for t in (n.mod.bn.running_mean, n.mod.bn.running_var, n.mod.bn.num_batches_tracked))

    def run_forward(self, fx: torch.Tensor, fy: torch.Tensor) -> torch.Tensor:
        """``forward`` - eagerly, or (``use_graph``) as a replay of its captured launch list."""
        if not self.use_graph:
            return self.forward(fx, fy)
        key = self._storage_key()
        if self._graph_key != key:
            self._graphs = [None, None]
            self._eager_calls = [0, 0]
            self._graph_key = key
        if self._graphs[0] is None:
            if self._eager_calls[0] < 2:
                self._eager_calls[0] += 1
                return self.forward(fx, fy)
            self._fx_in, self._fy_in = fx.clone(), fy.clone()
            torch.cuda.synchronize(self.device)
            gen = self.generation
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, capture_error_mode="thread_local"):
                    self._disp_out = self.forward(self._fx_in, self._fy_in)
            except RuntimeError as e:            # capture refused: keep launching eagerly, loudly
                return self._graph_failed(e, lambda: self.forward(fx, fy), gen)
            self.generation = gen
            self._graphs[0] = g
        self.generation += 1
        self._fx_in.copy_(fx)
        self._fy_in.copy_(fy)
        self._graphs[0].replay()
        return self._disp_out.clone()

    def _graph_failed(self, e, eager, generation):
        import warnings
        warnings.warn("leastereo_b200: CUDA graph capture of the training plan failed (%s); launching eagerly" % str(e)[:200])
        self.use_graph = False
        self._graphs = [None, None]
        self._no_graph = True
        torch.cuda.synchronize(self.device)
        self.generation = generation
        return eager()

    def run_backward(self, gdisp: torch.Tensor):
        if not self.use_graph or self._graphs[0] is None:
            return self.backward(gdisp)
        if self._graphs[1] is None:
            if self._eager_calls[1] < 1:
                self._eager_calls[1] += 1
                return self.backward(gdisp)
            self._gdisp_in = gdisp.contiguous().float().clone()
            torch.cuda.synchronize(self.device)
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, capture_error_mode="thread_local"):
                    self._bwd_out = self.backward(self._gdisp_in)
            except RuntimeError as e:
                return self._graph_failed(e, lambda: self.backward(gdisp), self.generation)
            self._graphs[1] = g
        self._gdisp_in.copy_(gdisp)
        self._graphs[1].replay()
        return self._bwd_out

    def backward(self, gdisp: torch.Tensor):
        self.pflat.zero_()
        for g in self.zero_grads:                  # only the volumes the first-writer analysis could not cover
            g.t.zero_()
        self.dmat = self.ops.disp_head_bwd(self.mat, gdisp.contiguous().float(), self.maxdisp)
        for n in reversed(self.nodes):
            n.backward()
        dfx, dfy = self.ops.cost_volume_bwd(self.grad_of(self.cost), self.m.initial_fm)
        return dfx, dfy, self.param_grads


class _ConvBRNode(_Node):
    """Conv3d -> BatchNorm3d(train) -> ReLU (flags per module), output optionally accumulated (state sums).
    ``upsample_to``: the (1x1x1) conv runs on the low-resolution input and its raw output is interpolated to that size
    before the batch statistics (TrainPlan._pre)."""

    def __init__(self, plan: TrainPlan, name, mod: ConvBR3d, src: Slice, dst: Slice, accumulate: bool, upsample_to=None):
        self.plan, self.name, self.mod, self.src, self.dst, self.accumulate = plan, name, mod, src, dst, accumulate
        c_out = mod.conv.out_channels
        if not mod.use_bn:
            raise LeaError("%s: ConvBR without BN is only supported as the last_3 head" % name)
        self.up = upsample_to is not None
        self.xl = Slice(plan.vol(c_out, src.spatial), 0, c_out) if self.up else None      # raw conv output, low-res
        self.x = Slice(plan.vol(c_out, upsample_to if self.up else src.spatial), 0, c_out)  # raw output BN sees (kept)
        dev = plan.device
        self.dx = Slice(PlanesVol.empty(self.x.vol.B, c_out, plan.P, *self.x.spatial, dev), 0, c_out)
        self.dxl = Slice(PlanesVol.empty(self.x.vol.B, c_out, plan.P, *src.spatial, dev), 0, c_out) if self.up else None
        self.dgrad_accumulate = True
        self.img_f: list = []
        self.img_b: list = []
        # per-channel BN vectors and the chunk partials of the reductions: allocated once, filled by kernels
        self.partial = torch.empty((max(64, 1184 // max(1, c_out >> 3)), 2, c_out), dtype=torch.float32, device=dev)
        self.mean, self.invstd, self.scale, self.shift, self.ka, self.kb, self.kc = (
            torch.empty(c_out, dtype=torch.float32, device=dev) for _ in range(7))

    def plan_backward(self, cover: _Coverage):
        cover.read(self.dst)
        self.dgrad_accumulate = cover.write(self.src)

    def forward(self):
        plan, mod, ops = self.plan, self.mod, self.plan.ops
        c = mod.conv.out_channels
        if self.up:
            plan.conv_raw(self.src, mod.conv.weight.detach(), self.xl, images=self.img_f)
            ops.trilinear_ac(self.xl.vol, 0, c, self.x.vol, 0)
        else:
            plan.conv_raw(self.src, mod.conv.weight.detach(), self.x, images=self.img_f)
        n = float(self.x.vol.B * _prod(self.x.spatial))
        bn = mod.bn
        # batch statistics: chunked per-channel sums, then ONE kernel for mean / invstd / scale / shift and the running
        # statistics (momentum, unbiased variance, num_batches_tracked) - no host arithmetic, no tiny tensor ops
        chunks = ops.channel_reduce_partial(self.x.vol, self.x.c0, c, self.partial, mode=0)
        track = bn.track_running_stats and bn.running_mean is not None
        ops.bn_finalize(self.partial, chunks, c, n, bn.weight.detach(), bn.bias.detach(), bn.eps,
                        bn.momentum if bn.momentum is not None else 0.1,
                        bn.running_mean if track else None, bn.running_var if track else None,
                        bn.num_batches_tracked if track else None, self.mean, self.invstd, self.scale, self.shift)
        ops.affine_relu(self.x.vol, self.x.c0, self.dst.vol, self.dst.c0, c, self.scale, self.shift, mod.relu,
                        self.accumulate)

    def backward(self):
        plan, mod, ops = self.plan, self.mod, self.plan.ops
        c = mod.conv.out_channels
        k = mod.conv.weight.shape[2]
        dy = plan.grad_of(self.dst.vol)
        n = float(self.x.vol.B * _prod(self.x.spatial))
        chunks = ops.channel_reduce_partial(self.x.vol, self.x.c0, c, self.partial, mode=1, dy=dy, dy_c0=self.dst.c0,
                                            relu=mod.relu, scale=self.scale, shift=self.shift, mean=self.mean,
                                            invstd=self.invstd)
        ops.bn_bwd_coeffs(self.partial, chunks, c, n, mod.bn.weight.detach(), self.invstd, self.ka, self.kb, self.kc,
                          plan.param_grads.get(mod.bn.weight), plan.param_grads.get(mod.bn.bias), accumulate=True)
        ops.bn_relu_bwd(self.x.vol, self.x.c0, dy, self.dst.c0, self.dx.vol, 0, c, mod.relu, self.scale, self.shift,
                        self.mean, self.invstd, self.ka, self.kb, self.kc)
        dconv = self.dx
        if self.up:                                # through the interpolation: gradient of the low-res raw output
            ops.trilinear_ac_bwd(self.dx.vol, 0, self.dxl.vol, 0, c, accumulate=False)
            dconv = self.dxl
        w = mod.conv.weight.detach()
        dw = plan.param_grads.get(mod.conv.weight)
        if dw is not None:
            ops.conv3d_wgrad(self.src.vol, self.src.c0, self.src.c, dconv.vol, 0, c, k, dw,
                             tensor_cores=(plan.conv_mode == "tc"))
        # data gradient: conv with transposed, tap-flipped weights into the input's gradient (first writer overwrites)
        dsrc = plan.grad_of(self.src.vol)
        plan.conv_raw(dconv, None, Slice(dsrc, self.src.c0, self.src.c), accumulate=self.dgrad_accumulate,
                      dgrad_of=w.contiguous(), images=self.img_b)


class _Last3Node(_Node):
    """``last_3``: Conv 32->1, no BN, no ReLU (skip_model_3d.py:132); fp32 output ``mat``."""

    def __init__(self, plan: TrainPlan, name, mod: ConvBR3d, src: Slice):
        self.plan, self.name, self.mod, self.src = plan, name, mod, src
        self.dpad = None
        self.dgrad_accumulate = True
        B, (D, H, W) = plan.B, plan.spatial
        self.dm8 = torch.zeros((B, 8, D, H, W), dtype=torch.float32, device=plan.device)   # channel 0 = dmat, rest 0
        w = mod.conv.weight
        self.wt = torch.zeros((w.shape[1], 8, 3, 3, 3), dtype=torch.float32, device=plan.device)
        # weight gradient on the tensor-core kernel with the output gradient padded to 8 channels (row 0 is last_3's)
        self.wg_tc = (plan.conv_mode == "tc" and plan.ops.device_build and plan.P == 2 and
                      bool(plan.ops.lib.lea_conv3d_wgrad_tc_supported(int(w.shape[1]), 8, 3, plan.P)))
        self.dw8 = torch.zeros((8,) + tuple(w.shape[1:]), dtype=torch.float32, device=plan.device) if self.wg_tc else None

    def plan_backward(self, cover: _Coverage):
        self.dgrad_accumulate = cover.write(self.src)

    def forward(self):
        self.plan.conv_raw(self.src, self.mod.conv.weight.detach(), None, dst_f32=self.plan.mat)

    def backward(self):
        plan, mod, ops = self.plan, self.mod, self.plan.ops
        w = mod.conv.weight.detach()                                  # (1, C, 3, 3, 3)
        self.dm8[:, 0].copy_(plan.dmat.reshape(self.dm8[:, 0].shape))
        self.dpad = ops.pack(self.dm8, plan.P, out=self.dpad)
        dw = plan.param_grads.get(mod.conv.weight)
        if dw is not None and self.wg_tc:
            self.dw8.zero_()
            ops.conv3d_wgrad(self.src.vol, self.src.c0, self.src.c, self.dpad, 0, 8, 3, self.dw8, tensor_cores=True)
            dw.add_(self.dw8[0:1])
        elif dw is not None:
            ops.conv3d_wgrad(self.src.vol, self.src.c0, self.src.c, self.dpad, 0, 1, 3, dw)
        self.wt[:, 0].copy_(w.flip(2, 3, 4)[0])
        dsrc = plan.grad_of(self.src.vol)
        plan.conv_raw(Slice(self.dpad, 0, 8), self.wt, Slice(dsrc, self.src.c0, self.src.c),
                      accumulate=self.dgrad_accumulate)


class _ResampleNode(_Node):
    def __init__(self, plan: TrainPlan, name, src: Slice, dst: Slice):
        self.plan, self.name, self.src, self.dst = plan, name, src, dst
        self.accumulate = True

    def plan_backward(self, cover: _Coverage):
        cover.read(self.dst)
        self.accumulate = cover.write(self.src)

    def forward(self):
        self.plan.ops.trilinear_ac(self.src.vol, self.src.c0, self.src.c, self.dst.vol, self.dst.c0)

    def backward(self):
        plan = self.plan
        plan.ops.trilinear_ac_bwd(plan.grad_of(self.dst.vol), self.dst.c0, plan.grad_of(self.src.vol), self.src.c0,
                                  self.src.c, accumulate=self.accumulate)


class _CopyNode(_Node):
    """``skip_connect`` (Identity) inside a step sum, or a state copy into the concat buffer."""

    def __init__(self, plan: TrainPlan, name, src: Slice, dst: Slice, accumulate: bool):
        self.plan, self.name, self.src, self.dst, self.accumulate = plan, name, src, dst, accumulate
        self.bwd_accumulate = True

    def plan_backward(self, cover: _Coverage):
        cover.read(self.dst)
        self.bwd_accumulate = cover.write(self.src)

    def forward(self):
        self.plan.ops.affine_relu(self.src.vol, self.src.c0, self.dst.vol, self.dst.c0, self.src.c, None, None, False,
                                  self.accumulate)

    def backward(self):
        plan = self.plan
        plan.ops.affine_relu(plan.grad_of(self.dst.vol), self.dst.c0, plan.grad_of(self.src.vol), self.src.c0,
                             self.src.c, None, None, False, self.bwd_accumulate)


# ------------------------------------------------------------------------------------------------------------
# autograd glue
# ------------------------------------------------------------------------------------------------------------

def used_parameters(plan: "TrainPlan") -> List[torch.nn.Parameter]:
    """Parameters that take part in this plan's forward (the reference leaves last_12/last_24/last_3.bn without
    gradients when the net ends on level 1)."""
    out: List[torch.nn.Parameter] = []
    seen = set()
    for node in plan.nodes:
        mod = getattr(node, "mod", None)
        if mod is None:
            continue
        cands = [mod.conv.weight] + ([mod.bn.weight, mod.bn.bias] if (mod.use_bn and isinstance(node, _ConvBRNode)) else [])
        for p in cands:
            if p.requires_grad and id(p) not in seen:
                seen.add(id(p))
                out.append(p)
    return out


class _HotPathTrainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan: TrainPlan, fx, fy, *params):
        ctx.plan = plan
        ctx.params = params
        with torch.no_grad():
            disp = plan.run_forward(fx.detach().float().contiguous(), fy.detach().float().contiguous())
        ctx.generation = plan.generation
        return disp

    @staticmethod
    def backward(ctx, gdisp):
        plan = ctx.plan
        if ctx.generation != plan.generation:
            # the single cached plan holds the saved activations (33 GB at 288x576 batch 4): a second train-mode
            # forward has overwritten the ones this backward needs
            raise RuntimeError("leastereo_b200: backward() of a train-mode forward whose saved activations were "
                               "overwritten by a later forward through the same module (one forward may be "
                               "outstanding per module; call backward() before the next train-mode forward)")
        with torch.no_grad():
            dfx, dfy, pg = plan.run_backward(gdisp)
        grads = []
        for p in ctx.params:
            g = pg.get(p)
            grads.append(None if g is None else g.reshape(p.shape).to(p.dtype))
        return (None, dfx, dfy) + tuple(grads)


_TRAIN_PLANS: Dict[tuple, TrainPlan] = {}


def hot_path_train_forward(model, fx: torch.Tensor, fy: torch.Tensor, ops: Optional[Ops] = None) -> torch.Tensor:
    """Train-mode hot path with autograd support: disparity (B, 3*H3, 3*W3) whose backward fills the gradients of the
    matching net's parameters and flows into the feature maps."""
    ops = ops or get_ops()
    opt = dict(getattr(model, "engine_options", None) or {})
    planes = int(opt.get("train_planes", opt.get("planes", 2)))
    conv = opt.get("conv", "tc") if ops.device_build else "simt"
    B, Cn, H3, W3 = fx.shape
    D3 = int(model.maxdisp / 3)
    key = (id(model.matching), str(fx.device), B, D3, H3, W3, planes, conv, id(ops))
    plan = _TRAIN_PLANS.get(key)
    if plan is not None and plan.m is not model.matching:       # id() of a collected module re-used by a new one
        plan = None
    if plan is None:
        _TRAIN_PLANS.clear()                       # one live training plan: buffers are large
        plan = TrainPlan(model.matching, ops, B, (D3, H3, W3), planes, fx.device, conv, model.maxdisp)
        _TRAIN_PLANS[key] = plan
    plan.use_graph = (bool(opt.get("train_graph", True)) and ops.device_build and fx.is_cuda
                      and not getattr(plan, "_no_graph", False))
    params = used_parameters(plan)
    # the BN kernels update running_mean / running_var / num_batches_tracked through raw pointers (no autograd version
    # bump): tell the eval-mode plans that parameters changed
    from .engine import bump_param_generation
    bump_param_generation()
    return _HotPathTrainFn.apply(plan, fx, fy, *params)
