"""Execution engine of the hot path: compiles the genotype-driven matching net into a flat list of kernel launches
over pre-allocated planes volumes, and runs cost volume -> matching net -> disparity head through the C-ABI library.

Dataflow being executed (reference lines): ``retrain/LEAStereo.py:34-51`` (cost volume, matching, disp),
``retrain/skip_model_3d.py:140-174`` (stems, 12 cells, the two skip concats + conv1/conv2, head),
``retrain/skip_model_3d.py:41-75`` (cell: resamples, 1x1x1 pre-processing, step sums, concat).

B200-first choices:
  * activations never leave the planes layout between layers; ``torch.cat`` becomes "write into a channel slice",
    the state sums become the conv epilogue's residual add, BN(eval)+ReLU live in the conv epilogue;
  * cells 1, 4 and 8 write their outputs side by side into ONE 3x-wide buffer so that both skip concats
    (``cat(C1,C4)``, ``cat(C4,C8)``) are plain channel slices of it - no copy;
  * every buffer, BN scale/shift vector and packed weight image has a fixed address for the life of a plan, so a
    whole forward can be captured in a CUDA graph.
"""
from __future__ import annotations

import threading
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch

from .kernels import LeaError, Ops, PlanesVol, get_ops, lea_tc_opts
from .modules import ConvBR2d, ConvBR3d, Identity2d, Identity3d, newFeature, newMatching

_IDENTITY = (Identity2d, Identity3d)
_CONVBR = (ConvBR2d, ConvBR3d)
from .structure import scale_dimension

_LOCK = threading.Lock()

# Bumped by anything that rewrites parameters behind autograd's back (``pipeline.FlatAdam.step`` updates the flat
# buffer through a raw pointer, which changes neither ``data_ptr`` nor ``_version`` of the parameters): part of every
# plan's parameter key, so packed weight images / BN vectors are rebuilt on the next eval-mode forward.
_PARAM_GENERATION = [0]


def bump_param_generation():
    _PARAM_GENERATION[0] += 1



@dataclass
class Slice:
    vol: PlanesVol
    c0: int
    c: int

    @property
    def spatial(self):
        return self.vol.spatial


@dataclass
class Step:
    kind: str                      # "conv_simt" | "conv_tc" | "resample"
    name: str                      # module path(s), for per-kernel timing reports
    flops: float = 0.0             # 2*M*N*K for convs, as launched
    bytes: float = 0.0             # algorithmic bytes read+written
    # conv
    p: object = None               # lea_conv
    mods: tuple = ()               # ConvBR3d modules whose weights are concatenated along c_out (batched conv)
    weight: Optional[torch.Tensor] = None     # fp32 (c_out, c_in, k, k, k) the launch reads (simt) / packs (tc)
    wcat: Optional[torch.Tensor] = None       # persistent concat buffer when len(mods) > 1
    opts: object = None
    image: Optional[torch.Tensor] = None      # packed tcgen05 weight image
    ref: Optional[torch.Tensor] = None
    wfn: object = None             # callable(step): fills step.wcat with weights DERIVED from the modules' parameters
    planes: int = 0                # plane count of the packed weight image (0 = the plan's)
    ref_flops: float = -1.0        # 2*M*N*K of the reference conv(s) this launch stands for (-1: same as flops)
    vols: tuple = ()               # every PlanesVol the launch reads or writes (buffer-liveness analysis)
    # resample
    rs: tuple = ()


class MatchingPlan:
    """Launch list + buffers of ``newMatching.forward`` for one (device, B, D, H, W, planes, conv mode)."""

    def __init__(self, matching: newMatching, ops: Ops, B: int, spatial: Tuple[int, int, int], planes: int,
                 device, conv_mode: str = "simt", mma_terms: int = 0, fuse: bool = True, tc_knobs: Optional[dict] = None):
        tc_knobs = tc_knobs or {}
        if conv_mode not in ("simt", "tc"):
            raise ValueError("conv mode must be 'simt' or 'tc'")
        self.m = matching
        self.ops = ops
        self.B, self.spatial, self.P = B, tuple(int(v) for v in spatial), planes
        self.device = torch.device(device)
        self.conv_mode = conv_mode
        self.mma_terms = mma_terms
        self.fuse = fuse               # graph-level rewrites: batched sibling convs, conv-before-upsample
        self.fuse_cv = bool(tc_knobs.get("fuse_cv", True))
        self.fuse_head = bool(tc_knobs.get("fuse_head", True))
        self.collapse_stem0 = bool(tc_knobs.get("collapse_stem0", True))
        self.fuse_resample_conv = bool(tc_knobs.get("fuse_resample_conv", False))
        self.accum_split = int(tc_knobs.get("accum_split", 0))
        self.acc_sets = int(tc_knobs.get("acc_sets", 0))
        self.tile_w_log2 = int(tc_knobs.get("tile_w_log2", 0))
        self.resident_weights = int(tc_knobs.get("resident_weights", 0))
        self.depth_chunk = int(tc_knobs.get("depth_chunk", 0))
        self.steps: List[Step] = []
        self.volumes: List[PlanesVol] = []
        self._bn_users: List[Tuple[ConvBR3d, int]] = []     # (module, offset into the BN buffers)
        self._bn_channels = 0
        self._resampled: Dict[tuple, Slice] = {}
        self.flat2d = False
        self._param_key = None
        # a module can have several launches (collapsed stem0: band launch + assemble), each with its own BN slice
        total_bn = 2 * sum(mod.conv.out_channels for mod in matching.modules() if isinstance(mod, _CONVBR)) + 256
        self.bn_scale = torch.ones(total_bn, dtype=torch.float32, device=self.device)
        self.bn_shift = torch.zeros(total_bn, dtype=torch.float32, device=self.device)
        self.reuse_buffers = bool(tc_knobs.get("reuse_buffers", True))
        self._arena: Optional[torch.Tensor] = None        # one allocation the activation volumes are views of
        self._arena_offsets: Optional[List[int]] = None   # byte offset of the i-th _vol() call inside the arena
        self._vol_calls = 0
        self._build()
        if self.reuse_buffers and self.volumes:
            self._rebuild_with_reuse()

    # ---- allocation helpers ---------------------------------------------------------------------------
    def _vol(self, c: int, spatial) -> PlanesVol:
        i = self._vol_calls
        self._vol_calls += 1
        if self._arena is not None:
            shape = (self.B, c // 8, self.P) + tuple(int(x) for x in spatial) + (8,)
            n = 2
            for x in shape:
                n *= x
            off = self._arena_offsets[i]
            v = PlanesVol(self._arena[off: off + n].view(torch.bfloat16).view(shape))
        else:
            # with buffer reuse the first build is a dry run: only sizes and the launch order matter, no memory is taken
            v = PlanesVol.empty(self.B, c, self.P, *spatial, "meta" if self.reuse_buffers else self.device)
        v._arena_index = i
        self.volumes.append(v)
        return v

    # ---- buffer reuse -----------------------------------------------------------------------------------
    def _rebuild_with_reuse(self):
        """The activation volumes of a forward pass are live only between their first writer and their last reader.
        Build 1 (plain allocations) records which launch touches which volume; the volumes are then packed into one
        arena by interval (first-fit over the launch order) and the plan is built a second time on views of it:
        same launch list, a fraction of the memory (KITTI: ~10 GB per pair -> see DESIGN.md)."""
        n = self._vol_calls
        first, last = [None] * n, [None] * n
        size = [0] * n
        for v in self.volumes:
            i = getattr(v, "_arena_index", None)
            if i is not None:
                size[i] = (v.nbytes() + 255) & ~255
        for k, st in enumerate(self.steps):
            for v in st.vols:
                i = getattr(v, "_arena_index", None)
                if i is None:
                    continue
                first[i] = k if first[i] is None else first[i]
                last[i] = k
        order = sorted(range(n), key=lambda i: (first[i] if first[i] is not None else 0))
        free: List[Tuple[int, int]] = []          # (offset, size) holes, sorted by offset
        live: List[Tuple[int, int, int]] = []     # (last step, offset, size)
        offsets = [0] * n
        top = peak = 0
        for i in order:
            f = first[i] if first[i] is not None else 0
            l = last[i] if last[i] is not None else len(self.steps)
            for item in [x for x in live if x[0] < f]:            # volumes dead before this one is first touched
                live.remove(item)
                free.append((item[1], item[2]))
            free.sort()
            merged: List[Tuple[int, int]] = []
            for off, sz in free:
                if merged and merged[-1][0] + merged[-1][1] == off:
                    merged[-1] = (merged[-1][0], merged[-1][1] + sz)
                else:
                    merged.append((off, sz))
            if merged and merged[-1][0] + merged[-1][1] == top:  # a hole at the end of the arena shrinks it
                top = merged[-1][0]
                merged.pop()
            free = merged
            slot = next((k for k, (off, sz) in enumerate(free) if sz >= size[i]), None)
            if slot is None:
                offsets[i] = top
                top += size[i]
            else:
                off, sz = free.pop(slot)
                offsets[i] = off
                if sz > size[i]:
                    free.append((off + size[i], sz - size[i]))
            live.append((l, offsets[i], size[i]))
            peak = max(peak, top)
        plain_bytes = sum(size)
        # tear down build 1, then build 2 on the arena
        self.steps, self.volumes = [], []
        self._bn_users, self._bn_channels = [], 0
        self._resampled = {}
        self._param_key = None
        self.cost = self.fxp = self.fyp = self.fxy = self.cv_maps = self.fxy3 = None
        if hasattr(self, "mat"):
            del self.mat
        self._vol_calls = 0
        self._arena_offsets = offsets
        self._arena = torch.empty(max(peak, 256), dtype=torch.uint8, device=self.device)
        self.plain_bytes, self.arena_bytes = plain_bytes, peak
        self._build()
        if self._vol_calls != n:
            raise LeaError("buffer reuse: the second build allocated %d volumes, the first %d" % (self._vol_calls, n))

    def workspace_bytes(self) -> int:
        if self._arena is not None:
            return int(self._arena.numel()) + sum(v.nbytes() for v in self.volumes if getattr(v, "_arena_index", None) is None)
        return sum(v.nbytes() for v in self.volumes)

    def _bn_slices(self, mods) -> Tuple[torch.Tensor, torch.Tensor]:
        """Contiguous BN scale/shift slices for a list of modules (concatenated in order)."""
        off0 = self._bn_channels
        for mod in mods:
            c = mod.conv.out_channels
            self._bn_users.append((mod, self._bn_channels))
            self._bn_channels += c
        return self.bn_scale[off0: self._bn_channels], self.bn_shift[off0: self._bn_channels]

    # ---- step emitters --------------------------------------------------------------------------------
    def _resample(self, name: str, src: Slice, spatial, dst: Optional[Slice] = None,
                  bn_of: Optional[ConvBR3d] = None, relu: bool = False) -> Slice:
        cache_key = None
        if dst is None and bn_of is None and not relu:
            # the same tensor is often resampled twice (as s1 of cell i and s0 of cell i+1): do it once
            cache_key = (id(src.vol), src.c0, src.c, tuple(spatial))
            if cache_key in self._resampled:
                return self._resampled[cache_key]
        if dst is None:
            dst = Slice(self._vol(src.c, spatial), 0, src.c)
        if cache_key is not None:
            self._resampled[cache_key] = dst
        scale = shift = None
        if bn_of is not None and bn_of.use_bn:
            scale, shift = self._bn_slices([bn_of])
        nbytes = 2.0 * self.P * src.c * self.B * (_prod(src.spatial) + _prod(spatial))
        self.steps.append(Step("resample", name, 0.0, nbytes,
                               rs=(src.vol, src.c0, src.c, dst.vol, dst.c0, scale, shift, relu), vols=(src.vol, dst.vol)))
        return dst

    def _emit_conv(self, name: str, mods, src: Slice, dst: Optional[Slice], *, res: bool = False,
                   dst_f32: Optional[torch.Tensor] = None, raw: bool = False,
                   fused_cv: bool = False):
        """Append one ConvBR launch.

        ``mods``: a ConvBR3d, or a list of ConvBR3d reading the same input whose outputs are adjacent channel
        slices (their weights / BN vectors are concatenated along c_out: one launch, one pass over the input).
        ``raw`` drops BN+ReLU (applied later by the up-sampling resample, see ``_resample``)."""
        wcat = None
        mods = tuple(mods) if isinstance(mods, (list, tuple)) else (mods,)
        w0 = mods[0].conv.weight
        c_in, k = w0.shape[1], w0.shape[2]
        is2d = w0.dim() == 4                   # 2-D conv run as a k^3 conv on a depth-1 volume (zero kd != 1 taps)
        c_out = sum(m.conv.out_channels for m in mods)
        relu = mods[0].relu and not raw
        use_bn = mods[0].use_bn and not raw
        for m in mods[1:]:
            if (m.conv.weight.shape[1], m.conv.weight.shape[2], m.relu, m.use_bn) != (c_in, k, mods[0].relu, mods[0].use_bn):
                raise LeaError("%s: batched convs must agree on c_in, kernel size, bn and relu" % name)
        scale = shift = None
        if use_bn:
            scale, shift = self._bn_slices(mods)
        if len(mods) == 1 and not is2d:
            weight = w0.detach()
        else:
            wcat = torch.zeros((c_out, c_in, k, k, k), dtype=torch.float32, device=self.device)
            weight = wcat
        if src.c != c_in:
            raise LeaError("%s: input has %d channels, conv expects %d" % (name, src.c, c_in))
        out_spatial = dst.spatial if (fused_cv and dst is not None) else src.spatial
        if dst is not None and dst.c != c_out:
            raise LeaError("%s: output slice has %d channels, conv produces %d" % (name, dst.c, c_out))
        p = self.ops.make_conv(src.vol, src.c0, c_in, c_out, k, scale, shift, relu,
                               dst=None if dst is None else dst.vol, dst_c0=0 if dst is None else dst.c0,
                               res=dst.vol if res else None, res_c0=dst.c0 if res else 0, dst_f32=dst_f32)
        m_vox = self.B * _prod(out_spatial)
        flops = 2.0 * m_vox * c_out * c_in * k ** 3
        if fused_cv:
            nbytes = 2.0 * self.P * self.B * (_prod(src.spatial) * c_in + _prod(out_spatial) * c_out)
        elif dst_f32 is None:
            nbytes = 2.0 * self.P * m_vox * (c_in + c_out * (2 if res else 1))
        else:
            nbytes = 2.0 * self.P * m_vox * c_in + 4.0 * m_vox * c_out
        use_tc = self.conv_mode == "tc" and \
            self.ops.tc_weight_image_bytes(c_in, c_out, k, self.P) > 0
        opts = None
        if use_tc:
            opts = lea_tc_opts()
            opts.mma_terms = self.mma_terms
            if fused_cv:
                opts.fused_cv = 1
                opts.fx, opts.fy = self.fxp.struct(), self.fyp.struct()
                opts.d3 = out_spatial[0]
                opts.cv_maps = self.cv_maps.data_ptr()
            opts.accum_split = self.accum_split
            opts.acc_sets = self.acc_sets
            opts.tile_w_log2 = self.tile_w_log2
            opts.resident_weights = self.resident_weights
            opts.depth_chunk = self.depth_chunk
        self.steps.append(Step("conv_tc" if use_tc else "conv_simt", name, flops, nbytes, p=p, mods=mods,
                               weight=weight, wcat=wcat, opts=opts, ref=src.vol.t,
                               vols=(src.vol,) + ((dst.vol,) if dst is not None else ())))

    def _emit_copy(self, name: str, src: Slice, dst: Slice, accumulate: bool):
        """``Identity`` (``skip_connect``, operations_3d.py:84-90) inside a step sum, or a state copied into its concat
        slot: ``dst (+)= src`` as one elementwise pass (``lea_affine_relu`` without scale / ReLU) - no weights, no MMA."""
        if src.c != dst.c or src.spatial != dst.spatial:
            raise LeaError("%s: identity needs matching shapes" % name)
        nbytes = 2.0 * self.P * self.B * _prod(src.spatial) * src.c * (3 if accumulate else 2)
        self.steps.append(Step("copy", name, 0.0, nbytes, rs=(src, dst, bool(accumulate)), vols=(src.vol, dst.vol)))

    def _emit_tap_projection(self, name: str, mod: ConvBR3d, src: Slice, dst: Slice):
        """1x1x1 conv C -> k^3 "tap" channels with weights W[t, c] = mod.weight[0, c, kd, kh, kw], t = kd*9+kh*3+kw:
        the channel contraction of a single-output 3x3x3 conv, run BEFORE the up-sample it follows
        (skip_model_3d.py:162-169; see ``lea_head_taps``).  Rows >= k^3 of the padded weight stay zero."""
        w3 = mod.conv.weight
        c_in, taps = w3.shape[1], w3.shape[2] * w3.shape[3] * w3.shape[4]
        if w3.shape[0] != 1 or taps > dst.c or src.c != c_in:
            raise LeaError("%s: tap projection needs a single-output conv and %d <= %d tap channels" % (name, taps, dst.c))
        wcat = torch.zeros((dst.c, c_in, 1, 1, 1), dtype=torch.float32, device=self.device)

        def fill(step, mod=mod, taps=taps, c_in=c_in):
            step.wcat[:taps].copy_(mod.conv.weight.detach().reshape(c_in, taps).t().reshape(taps, c_in, 1, 1, 1))

        p = self.ops.make_conv(src.vol, src.c0, c_in, dst.c, 1, None, None, False, dst=dst.vol, dst_c0=dst.c0)
        m_vox = self.B * _prod(src.spatial)
        use_tc = self.conv_mode == "tc" and self.ops.tc_weight_image_bytes(c_in, dst.c, 1, self.P) > 0
        opts = None
        if use_tc:
            opts = lea_tc_opts()
            opts.mma_terms = self.mma_terms
            opts.accum_split = self.accum_split
            opts.acc_sets = self.acc_sets
            opts.tile_w_log2 = self.tile_w_log2
            opts.resident_weights = self.resident_weights
            opts.depth_chunk = self.depth_chunk
        self.steps.append(Step("conv_tc" if use_tc else "conv_simt", name, 2.0 * m_vox * taps * c_in,
                               2.0 * self.P * m_vox * (c_in + dst.c), p=p, mods=(mod,), weight=wcat, wcat=wcat, opts=opts,
                               ref=src.vol.t, wfn=fill, vols=(src.vol, dst.vol)))

    # ---- collapsed stem0 ------------------------------------------------------------------------------------
    def _can_collapse_stem0(self, fm: int, L0) -> bool:
        st = self.m.stem0
        w = st.conv.weight
        return (self.fuse and self.collapse_stem0 and self.conv_mode == "tc" and tuple(w.shape[2:]) == (3, 3, 3)
                and w.shape[1] == 2 * fm and w.shape[0] % 8 == 0 and w.shape[0] <= 32 and L0[0] >= 3 and L0[2] >= 24
                and self.ops.tc_weight_image_bytes(fm, 2 * w.shape[0], 3, 3) > 0)

    def _emit_collapsed_stem0(self, fm: int, L0, v0: PlanesVol):
        """stem0 on the cost volume (LEAStereo.py:34-48 + skip_model_3d.py:141) without the per-disparity work where
        the volume is un-masked: out = L[h,w] + A[h,w-d-1] + B[h,w-d+1] with three 2-D convs of the feature maps
        (see ``lea_stem0_assemble``); the tensor-core kernel with the fused loader does the remaining voxels."""
        m, B = self.m, self.B
        st = m.stem0
        c_out = st.conv.out_channels
        D, H, W = L0
        P3 = 3                                               # the maps are summed afterwards: keep them exact
        self.fxy3 = PlanesVol.empty(2 * B, fm, P3, 1, H, W, self.device)
        fx3, fy3 = PlanesVol(self.fxy3.t[:B]), PlanesVol(self.fxy3.t[B:])
        lmap = PlanesVol.empty(B, c_out, P3, 1, H, W, self.device)
        abmap = PlanesVol.empty(B, 2 * c_out, P3, 1, H, W, self.device)
        self.volumes += [self.fxy3, lmap, abmap]

        def fill_l(step, st=st, fm=fm):
            step.wcat.copy_(collapsed_stem0_weights(st.conv.weight.detach(), fm)[0])

        def fill_ab(step, st=st, fm=fm):
            step.wcat.copy_(collapsed_stem0_weights(st.conv.weight.detach(), fm)[1])

        def emit_map(name, src_vol, dst_vol, co, fill):
            wcat = torch.zeros((co, fm, 3, 3, 3), dtype=torch.float32, device=self.device)
            p = self.ops.make_conv(src_vol, 0, fm, co, 3, None, None, False, dst=dst_vol, dst_c0=0)
            opts = lea_tc_opts()
            opts.accum_split = self.accum_split
            self.steps.append(Step("conv_tc", name, 2.0 * B * H * W * co * fm * 9,
                                   2.0 * P3 * B * H * W * (fm + co), p=p, mods=(st,), weight=wcat, wcat=wcat, opts=opts,
                                   ref=src_vol.t, wfn=fill, planes=P3, vols=(src_vol, dst_vol)))

        emit_map("stem0.collapsed.L(2-D)", fx3, lmap, c_out, fill_l)
        emit_map("stem0.collapsed.A|B(2-D)", fy3, abmap, 2 * c_out, fill_ab)
        # the voxels the maps do not cover: tensor-core kernel with the fused cost-volume loader
        n_before = len(self.steps)
        self._emit_conv("stem0(fused cost volume; band + edges)", st, Slice(self.fxp, 0, 2 * fm), Slice(v0, 0, fm),
                        fused_cv=True)
        tc_step = self.steps[n_before]
        tc_step.opts.cv_skip = 1
        tiles_w = (W + 7) // 8
        interior = sum(1 for d in range(D) for tw in range(tiles_w)
                       if (d >= 1 and d <= D - 2 and 8 * tw >= d + 2 and 8 * tw + 7 <= W - 2) or 8 * tw + 7 <= d - 3)
        frac = 1.0 - interior * 8.0 / (D * W)
        tc_step.ref_flops = tc_step.flops             # the reference's stem0: the whole volume
        tc_step.flops *= frac
        for ms in self.steps[n_before - 2: n_before]:
            ms.ref_flops = 0.0                         # the 2-D maps have no counterpart of their own
        scale, shift = self._bn_slices([st]) if st.use_bn else (None, None)
        nbytes = 2.0 * self.P * B * D * H * W * c_out * (1.0 - frac) + 2.0 * P3 * B * H * W * 3 * c_out
        self.steps.append(Step("stem0_assemble", "stem0.collapsed.assemble", 0.0, nbytes,
                               rs=(lmap, abmap, v0, 0, c_out, scale, shift, st.relu), vols=(lmap, abmap, v0)))

    def _max_batched_c_out(self) -> int:
        return 64

    # ---- network construction -------------------------------------------------------------------------
    def _pre(self, name: str, mod: ConvBR3d, src: Slice, spatial, dst: Slice):
        """Cell input path: [resample to ``spatial``] -> 1x1x1 ConvBR -> ``dst``  (skip_model_3d.py:44-53).
        When the resample ENLARGES the volume the conv runs first on the small volume (it commutes with the
        interpolation) and BN+ReLU are applied by the resample kernel's epilogue."""
        if src.spatial == tuple(spatial):
            self._emit_conv(name, mod, src, dst)
        elif self.fuse and _prod(src.spatial) < _prod(spatial):
            small = Slice(self._vol(dst.c, src.spatial), 0, dst.c)
            self._emit_conv(name + "(raw,low-res)", mod, src, small, raw=True)
            self.steps[-1].ref_flops = self.steps[-1].flops * _prod(spatial) / _prod(src.spatial)   # reference: after the up-sample
            self._resample(name + ".upsample+bn+relu", small, spatial, dst=dst, bn_of=mod, relu=mod.relu)
        elif self.fuse and self.fuse_resample_conv and self._can_fuse_resample_conv(mod, src, dst):
            self._emit_resample_conv(name + "(fused into the down-sampling resample)", mod, src, dst)
        else:
            r = self._resample(name + ".resample", src, spatial)
            self._emit_conv(name, mod, r, dst)

    def _can_fuse_resample_conv(self, mod: ConvBR3d, src: Slice, dst: Slice) -> bool:
        w = mod.conv.weight
        return (all(int(k) == 1 for k in w.shape[2:]) and w.shape[0] == dst.c and w.shape[1] == src.c
                and dst.c % 8 == 0 and dst.c <= 64 and src.c % 8 == 0 and src.c <= 256)

    def _emit_resample_conv(self, name: str, mod: ConvBR3d, src: Slice, dst: Slice):
        """Down-sampling resample + the 1x1x1 ConvBR that consumes it in one pass over the source
        (skip_model_3d.py:44-53; ``lea_resample_conv1x1``): the resampled intermediate is never written."""
        scale, shift = self._bn_slices([mod]) if mod.use_bn else (None, None)
        m_vox = self.B * _prod(dst.spatial)
        nbytes = 2.0 * self.P * self.B * (_prod(src.spatial) * src.c + _prod(dst.spatial) * dst.c)
        self.steps.append(Step("resample_conv1x1", name, 2.0 * m_vox * dst.c * src.c, nbytes, mods=(mod,),
                               weight=mod.conv.weight.detach(), rs=(src, dst, scale, shift, mod.relu),
                               vols=(src.vol, dst.vol)))

    def _cell(self, i: int, s0: Slice, s1: Slice, out: Optional[Slice] = None, cell=None, bm=None,
              prefix: str = "") -> Tuple[Slice, Slice]:
        cell = cell if cell is not None else self.m.cells[i]
        spec = cell.spec
        name = "%scells.%d" % (prefix, i)
        prev_input = s1
        c_out = spec.c_out
        sp = s1.spatial
        if spec.downup_sample != 0:
            sp = tuple(scale_dimension(n, spec.scale) for n in s1.spatial)
            if self.flat2d:
                sp = (1,) + sp[1:]             # a 2-D net runs as depth-1 volumes: only H and W are resampled
        bm = bm if bm is not None else self.m._block_multiplier
        n_states = 2 + len(spec.steps)
        first_in_concat = n_states - bm
        if first_in_concat < 0:
            raise LeaError("block_multiplier larger than the number of cell states")
        if out is None:
            out = Slice(self._vol(bm * c_out, sp), 0, bm * c_out)
        elif out.spatial != sp or out.c != bm * c_out:
            raise LeaError("cell %d output %s x%d does not fit the skip-concat slot %s x%d (the reference's torch.cat "
                           "would fail too)" % (i, sp, bm * c_out, out.spatial, out.c))
        slots: Dict[int, Slice] = {}

        def state_slot(q: int) -> Slice:
            if q not in slots:
                pos = q - first_in_concat
                slots[q] = Slice(out.vol, out.c0 + pos * c_out, c_out) if pos >= 0 else \
                    Slice(self._vol(c_out, sp), 0, c_out)
            return slots[q]

        # ---- the two pre-processed inputs (states 0 and 1)
        if s0.c != c_out:
            self._pre(name + ".pre_preprocess", cell.pre_preprocess, s0, sp, state_slot(0))
        else:
            r = s0 if s0.spatial == sp else self._resample(name + ".resample_s0", s0, sp)
            if first_in_concat <= 0:
                self._emit_copy(name + ".s0_copy", r, state_slot(0), accumulate=False)
            else:
                slots[0] = r
        self._pre(name + ".preprocess", cell.preprocess, s1, sp, state_slot(1))

        # ---- step sums: state q = sum of ops over selected earlier states (skip_model_3d.py:57-72)
        consumers: Dict[int, List[Tuple[int, int]]] = {}        # source state -> [(target state, ops index)]
        pending: Dict[int, int] = {}                             # target state -> contributions still missing
        for k, step in enumerate(spec.steps):
            pending[2 + k] = len(step)
            for (j, opi) in step:
                consumers.setdefault(j, []).append((2 + k, opi))
        written: Dict[int, bool] = {}
        done = {0, 1}
        emitted = set()
        while len(emitted) < len(consumers):
            ready = [j for j in consumers if j in done and j not in emitted]
            if not ready:
                raise LeaError("cell %d: genotype wiring has a cycle" % i)
            # the source with most consumers first: its sibling convs can then all be first writers -> one launch
            j = max(ready, key=lambda q: (len(consumers[q]), -q))
            emitted.add(j)
            todo = sorted(consumers[j])
            while todo:
                tgt, opi = todo[0]
                op = cell._ops[opi]
                group = [(tgt, opi)]
                if self.fuse and not isinstance(op, _IDENTITY):
                    # extend with siblings: consecutive target slots, same first-writer status, all convs
                    for (t2, o2) in todo[1:]:
                        pt, _ = group[-1]
                        same_buf = (state_slot(t2).vol is state_slot(pt).vol and
                                    state_slot(t2).c0 == state_slot(pt).c0 + c_out)
                        if (t2 == pt + 1 and same_buf and written.get(t2, False) == written.get(tgt, False)
                                and not isinstance(cell._ops[o2], _IDENTITY)
                                and (len(group) + 1) * c_out <= self._max_batched_c_out()):
                            group.append((t2, o2))
                        else:
                            break
                todo = [x for x in todo if x not in group]
                first = state_slot(group[0][0])
                dst = Slice(first.vol, first.c0, c_out * len(group))
                res = written.get(tgt, False)
                if isinstance(op, _IDENTITY):
                    # skip_connect (operations_3d.py:84-90): the state is added into (or copied to) the target slot
                    self._emit_copy("%s._ops.%d(skip)" % (name, opi), state_slot(j), dst, accumulate=res)
                else:
                    mods = [cell._ops[o] for (_, o) in group]
                    nm = "%s._ops.%s" % (name, "+".join(str(o) for (_, o) in group))
                    self._emit_conv(nm, mods, state_slot(j), dst, res=res)
                for (t2, _) in group:
                    written[t2] = True
                    pending[t2] -= 1
                    if pending[t2] == 0:
                        done.add(t2)
        return prev_input, out

    def _build(self):
        m = self.m
        fm = m.initial_fm
        L0 = self.spatial
        v0, v1 = self._vol(fm, L0), self._vol(fm, L0)
        self.cost = self.fxp = self.fyp = self.fxy = self.cv_maps = self.fxy3 = None
        can_fuse_cv = (self.fuse_cv and self.conv_mode == "tc" and fm % 16 == 0 and L0[0] <= L0[2] and
                       self.ops.tc_weight_image_bytes(2 * fm, fm, 3, self.P) > 0)
        if can_fuse_cv:
            # retrain/LEAStereo.py:34-48 inside stem0's TMA loader: the 2C x D3 x H3 x W3 volume is never materialised
            # one allocation for both feature maps (left block, right block) so that the native feature net can
            # write its batched (left+right) output straight into the loader's operands
            self.fxy = PlanesVol.empty(2 * self.B, fm, self.P, 1, L0[1], L0[2], self.device)
            self.fxp = PlanesVol(self.fxy.t[: self.B])
            self.fyp = PlanesVol(self.fxy.t[self.B:])
            self.cv_maps = self.ops.build_fused_cv_maps(self.fxp, self.fyp, L0[0])
            if self._can_collapse_stem0(fm, L0):
                self._emit_collapsed_stem0(fm, L0, v0)
            else:
                self._emit_conv("stem0(fused cost volume)", m.stem0, Slice(self.fxp, 0, 2 * fm), Slice(v0, 0, fm),
                                fused_cv=True)
        else:
            self.cost = self._vol(2 * fm, L0)
            self._emit_conv("stem0", m.stem0, Slice(self.cost, 0, 2 * fm), Slice(v0, 0, fm))
        self._emit_conv("stem1", m.stem1, Slice(v0, 0, fm), Slice(v1, 0, fm))
        stem0, stem1 = Slice(v0, 0, fm), Slice(v1, 0, fm)

        out0 = self._cell(0, stem0, stem1)
        # cells 1, 4, 8 share one buffer: [C1 | C4 | C8]  (skip_model_3d.py:150,155 concats become slices)
        spec1 = m.cells[1].spec
        sp1 = out0[1].spatial if spec1.downup_sample == 0 else \
            tuple(scale_dimension(n, spec1.scale) for n in out0[1].spatial)
        cw = m._block_multiplier * spec1.c_out
        skip = self._vol(3 * cw, sp1)
        out1 = self._cell(1, out0[0], out0[1], Slice(skip, 0, cw))
        out2 = self._cell(2, out1[0], out1[1])
        out3 = self._cell(3, out2[0], out2[1])
        out4 = self._cell(4, out3[0], out3[1], Slice(skip, cw, cw))
        x5 = Slice(self._vol(m.conv1.conv.out_channels, sp1), 0, m.conv1.conv.out_channels)
        self._emit_conv("conv1", m.conv1, Slice(skip, 0, 2 * cw), x5)
        out5 = self._cell(5, out4[0], x5)
        out6 = self._cell(6, out5[0], out5[1])
        out7 = self._cell(7, out6[0], out6[1])
        out8 = self._cell(8, out7[0], out7[1], Slice(skip, 2 * cw, cw))
        x9 = Slice(self._vol(m.conv2.conv.out_channels, sp1), 0, m.conv2.conv.out_channels)
        self._emit_conv("conv2", m.conv2, Slice(skip, cw, 2 * cw), x9)
        out9 = self._cell(9, out8[0], x9)
        out10 = self._cell(10, out9[0], out9[1])
        out11 = self._cell(11, out10[0], out10[1])
        last = out11[1]

        # head (skip_model_3d.py:161-173)
        d, h, w = L0
        self.mat = torch.empty((self.B, 1, d, h, w), dtype=torch.float32, device=self.device)

        def conv_to(name, mod, src: Slice) -> Slice:
            dst = Slice(self._vol(mod.conv.out_channels, src.spatial), 0, mod.conv.out_channels)
            self._emit_conv(name, mod, src, dst)
            return dst

        # `pre` = input of the final up-sample to L0 (None when the net already ends on level 0)
        if last.spatial[1] == h:
            feat, pre = last, None
        elif last.spatial[1] == h // 2:
            pre = conv_to("last_6", m.last_6, last)
        elif last.spatial[1] == h // 4:
            t = self._resample("head.upsample_12", conv_to("last_12", m.last_12, last), (d // 2, h // 2, w // 2))
            pre = conv_to("last_6", m.last_6, t)
        elif last.spatial[1] == h // 8:
            t = self._resample("head.upsample_24", conv_to("last_24", m.last_24, last), (d // 4, h // 4, w // 4))
            t = self._resample("head.upsample_12", conv_to("last_12", m.last_12, t), (d // 2, h // 2, w // 2))
            pre = conv_to("last_6", m.last_6, t)
        else:
            raise LeaError("matching net ends on a level the reference head does not handle (H3=%d, last H=%d)"
                           % (h, last.spatial[1]))
        l3 = m.last_3
        fusable = (pre is not None and self.fuse and self.fuse_head and tuple(l3.conv.weight.shape[2:]) == (3, 3, 3)
                   and l3.conv.out_channels == 1 and not l3.use_bn and not l3.relu
                   and all(o >= 2 * i - 1 for o, i in zip(L0, pre.spatial)))
        if fusable:
            # upsample_6 -> last_3 without the up-sampled volume: last_3's channel contraction runs on the small
            # volume (27 tap channels), the separable interpolation + tap shifts are summed by lea_head_taps
            q = Slice(self._vol(32, pre.spatial), 0, 32)
            self._emit_tap_projection("last_3.taps(low-res)", l3, pre, q)
            self.steps[-1].ref_flops = 2.0 * self.B * _prod(L0) * l3.conv.in_channels * 27      # last_3 on the up-sampled volume
            ws = self.ops.head_taps_workspace(q.vol, L0)
            nbytes = 2.0 * self.P * self.B * _prod(pre.spatial) * 32 + 4.0 * self.B * _prod(L0)
            self.steps.append(Step("head_taps", "head.upsample_6+last_3", 0.0, nbytes, rs=(q.vol, 0, self.mat, ws),
                                   vols=(q.vol,)))
            return
        if pre is not None:
            feat = self._resample("head.upsample_6", pre, L0)
        if feat.spatial != L0:
            raise LeaError("head input %s does not match the cost volume %s" % (feat.spatial, L0))
        self._emit_conv("last_3", m.last_3, feat, None, dst_f32=self.mat)

    # ---- parameters -----------------------------------------------------------------------------------
    def _current_param_key(self):
        key = [_PARAM_GENERATION[0]]
        for mod, _ in self._bn_users:
            for t in (mod.bn.weight, mod.bn.bias, mod.bn.running_mean, mod.bn.running_var):
                key.append((t.data_ptr(), t._version))
        for s in self.steps:
            for mod in s.mods:
                t = mod.conv.weight
                key.append((t.data_ptr(), t._version))
        return tuple(key)

    def refresh_params(self, force: bool = False):
        """Recompute BN scale/shift, sibling-conv weight concats and packed tensor-core weight images when any
        parameter changed (checked through tensor versions / storage addresses)."""
        key = self._current_param_key()
        if not force and key == self._param_key:
            return
        with torch.no_grad():
            if self._bn_users:
                users = [m for m, _ in self._bn_users]
                g = torch.cat([m.bn.weight.detach().float().reshape(-1) for m in users])
                b = torch.cat([m.bn.bias.detach().float().reshape(-1) for m in users])
                mu = torch.cat([m.bn.running_mean.detach().float().reshape(-1) for m in users])
                var = torch.cat([m.bn.running_var.detach().float().reshape(-1) for m in users])
                eps = torch.cat([torch.full((m.bn.num_features,), float(m.bn.eps)) for m in users]).to(g.device)
                scale = g / torch.sqrt(var + eps)
                self.bn_scale[: scale.numel()].copy_(scale)
                self.bn_shift[: scale.numel()].copy_(b - mu * scale)
            for s in self.steps:
                if s.kind == "resample_conv1x1":
                    w = s.mods[0].conv.weight
                    if not (w.is_contiguous() and w.dtype == torch.float32 and w.device == self.device):
                        raise LeaError("%s: weights must be contiguous fp32 on %s" % (s.name, self.device))
                    s.weight = w.detach()
                if not s.kind.startswith("conv"):
                    continue
                if s.mods:
                    for mod in s.mods:
                        w = mod.conv.weight
                        if not (w.is_contiguous() and w.dtype == torch.float32 and w.device == self.device):
                            raise LeaError("%s: weights must be contiguous fp32 on %s" % (s.name, self.device))
                    if s.wfn is not None:
                        s.wfn(s)
                    elif s.wcat is not None:
                        ws = [mod.conv.weight.detach() for mod in s.mods]
                        if ws[0].dim() == 4:       # 2-D weights go into the middle depth slice of the k^3 kernel
                            s.wcat[:, :, s.wcat.shape[2] // 2].copy_(torch.cat(ws, dim=0))
                        else:
                            torch.cat(ws, dim=0, out=s.wcat)
                    else:
                        s.weight = s.mods[0].conv.weight.detach()
                if s.kind == "conv_tc":
                    s.image = self.ops.pack_weights_tc(s.weight, s.planes or self.P, out=s.image)
        self._param_key = key

    # ---- execution ------------------------------------------------------------------------------------
    def run_step(self, s: Step):
        if s.kind == "resample":
            src, c0, c, dst, dst_c0, scale, shift, relu = s.rs
            self.ops.trilinear_ac(src, c0, c, dst, dst_c0, scale, shift, relu)
        elif s.kind == "resample_conv1x1":
            src, dst, scale, shift, relu = s.rs
            self.ops.resample_conv1x1(src.vol, src.c0, src.c, [(dst.vol, dst.c0, dst.c, s.weight, scale, shift, relu)])
        elif s.kind == "conv_simt":
            self.ops.conv3d_simt(s.p, s.weight, s.ref)
        elif s.kind == "conv_tc":
            self.ops.conv3d_tc(s.p, s.image, s.opts, s.ref)
        elif s.kind == "stem0_assemble":
            lmap, abmap, dst, dst_c0, c_out, scale, shift, relu = s.rs
            self.ops.stem0_assemble(lmap, abmap, dst, dst_c0, c_out, scale, shift, relu)
        elif s.kind == "head_taps":
            q, q_c0, mat, ws = s.rs
            self.ops.head_taps(q, q_c0, mat, ws)
        elif s.kind == "copy":
            src, dst, accumulate = s.rs
            self.ops.affine_relu(src.vol, src.c0, dst.vol, dst.c0, src.c, None, None, False, accumulate)
        elif s.kind == "repack":
            src, dst, c = s.rs
            self.ops.affine_relu(src, 0, dst, 0, c, None, None, False, False)
        elif s.kind == "feature_stem":
            img, w0, sc0, sh0, w1, sc1, sh1, dst = s.rs
            self.ops.feature_stem(img, w0, sc0, sh0, w1, sc1, sh1, dst, 0)
        else:
            raise LeaError("unknown step " + s.kind)

    def run(self, check_params: bool = True) -> torch.Tensor:
        """Runs the launch list over ``self.cost`` (already filled) and returns ``self.mat``."""
        if check_params or self._param_key is None:
            self.refresh_params()
        for s in self.steps:
            self.run_step(s)
        return getattr(self, "mat", None)

    def conv_flops(self) -> float:
        return sum(s.flops for s in self.steps)

    def reference_conv_flops(self) -> float:
        """2*M*N*K summed over the reference's own conv list (1344.17 GFLOP per KITTI pair)."""
        return sum((s.ref_flops if s.ref_flops >= 0.0 else s.flops) for s in self.steps)


class FeaturePlan(MatchingPlan):
    """Launch list of the 2D feature net (``retrain/new_model_2d.py:129-165``) for a batch of left+right images.

    SURVEY.md §8(f) row 1.  The net runs on the same kernels as the 3D net: activations are depth-1 planes volumes,
    a 3x3 conv is a 3x3x3 conv whose kd != 1 taps are zero (only the kd = 1 MMA is issued at depth 1, so nothing is
    wasted), bilinear align_corners=True resampling is the trilinear kernel at depth 1.  stem0+stem1 (the only
    full-resolution / strided layers) are one fused CUDA-core kernel.  The last 1x1 conv writes directly into the
    operand buffer of the fused cost-volume loader (left block | right block)."""

    def __init__(self, feature: newFeature, ops: Ops, B2: int, H: int, W: int, planes: int, device, out: PlanesVol,
                 mma_terms: int = 0, fuse: bool = True, tc_knobs: Optional[dict] = None, conv_mode: str = "tc",
                 out3: Optional[PlanesVol] = None):
        self.img_hw = (int(H), int(W))
        self.out = out
        self.out3 = out3               # optional second destination holding the features in this plan's own planes
        h3, w3 = (H - 1) // 3 + 1, (W - 1) // 3 + 1
        super().__init__(feature, ops, B2, (1, h3, w3), planes, device, conv_mode, mma_terms, fuse, tc_knobs)

    def _build(self):
        f = self.m
        self.flat2d = True
        L0 = self.spatial
        H, W = self.img_hw
        fm = f._filter_multiplier * f._block_multiplier
        if f.stem1.conv.stride[0] != 3 or f.stem0.conv.out_channels > 16 or fm > 32 or \
                self.out.spatial != L0 or self.out.C != fm or self.out.B != self.B:
            raise LeaError("native feature net: unsupported stem geometry")
        self.img = torch.empty((self.B, 3, H, W), dtype=torch.float32, device=self.device)
        stem1 = Slice(self._vol(fm, L0), 0, fm)
        sc0, sh0 = self._bn_slices([f.stem0])
        sc1, sh1 = self._bn_slices([f.stem1])
        flops = 2.0 * self.B * (H * W * f.stem0.conv.out_channels * 27 + _prod(L0) * fm * f.stem0.conv.out_channels * 9)
        self.steps.append(Step("feature_stem", "feature.stem0+stem1", flops, 4.0 * self.B * 3 * H * W,
                               mods=(f.stem0, f.stem1), rs=(self.img, None, sc0, sh0, None, sc1, sh1, stem1.vol),
                               vols=(stem1.vol,)))
        stem2 = Slice(self._vol(fm, L0), 0, fm)
        self._emit_conv("feature.stem2", f.stem2, stem1, stem2)
        out = (stem1, stem2)
        for i, cell in enumerate(f.cells):
            out = self._cell(i, out[0], out[1], cell=cell, bm=f._block_multiplier, prefix="feature.")
        last = out[1]
        _, h, w = L0

        def conv_to(name, mod, src: Slice) -> Slice:
            dst = Slice(self._vol(mod.conv.out_channels, src.spatial), 0, mod.conv.out_channels)
            self._emit_conv(name, mod, src, dst)
            return dst

        if last.spatial[1] == h:
            feat = last
        elif last.spatial[1] == h // 2:
            feat = self._resample("feature.upsample_6", conv_to("feature.last_6", f.last_6, last), L0)
        elif last.spatial[1] == h // 4:
            t = self._resample("feature.upsample_12", conv_to("feature.last_12", f.last_12, last), (1, h // 2, w // 2))
            feat = self._resample("feature.upsample_6", conv_to("feature.last_6", f.last_6, t), L0)
        else:
            raise LeaError("feature net ends on a level the native path does not handle")
        used3 = False
        if self.out.P == self.P:
            self._emit_conv("feature.last_3", f.last_3, feat, Slice(self.out, 0, fm))
        else:
            # the feature net runs with its own plane count (3 = exact storage / bf16x6: it is tiny, and its rounding
            # errors are amplified by the whole matching net); convert to the matching net's operand format at the end
            if self.out3 is not None and self.out3.P == self.P and self.out3.spatial == L0 and self.out3.C == fm:
                tmp = Slice(self.out3, 0, fm)          # the collapsed stem0 reads the features at full precision
                used3 = True
            else:
                tmp = Slice(self._vol(fm, L0), 0, fm)
            self._emit_conv("feature.last_3", f.last_3, feat, tmp)
            self.steps.append(Step("repack", "feature.to_operand_planes", 0.0,
                                   2.0 * self.B * _prod(L0) * fm * (self.P + self.out.P), rs=(tmp.vol, self.out, fm),
                                   vols=(tmp.vol, self.out)))
        if self.out3 is not None and not used3:
            self.steps.append(Step("repack", "feature.to_map_planes", 0.0,
                                   2.0 * self.B * _prod(L0) * fm * (self.out.P + self.out3.P), rs=(self.out, self.out3, fm),
                                   vols=(self.out, self.out3)))

    def refresh_params(self, force: bool = False):
        key = self._current_param_key()
        if not force and key == self._param_key:
            return
        super().refresh_params(force=True)
        for s in self.steps:
            if s.kind == "feature_stem":
                img, _, sc0, sh0, _, sc1, sh1, dst = s.rs
                s.rs = (img, s.mods[0].conv.weight.detach().contiguous(), sc0, sh0,
                        s.mods[1].conv.weight.detach().contiguous(), sc1, sh1, dst)


def collapsed_stem0_weights(w: torch.Tensor, fm: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """2-D kernels of the collapsed stem0 from its (c_out, 2*fm, 3, 3, 3) weight, as k^3 kernels whose kd != 1 taps
    are zero (they run on depth-1 volumes):  L (c_out, fm, 3,3,3) = sum over kd of the x half;  A|B (2*c_out, fm, 3,3,3)
    from the y half, taps delta = kw - kd in {-2,-1,0} for A (read at u-1) and {1,2} for B (read at u+1)."""
    c_out = w.shape[0]
    wd = w.double()
    wl = torch.zeros((c_out, fm, 3, 3, 3), dtype=torch.float32, device=w.device)
    wl[:, :, 1].copy_(wd[:, :fm].sum(dim=2).float())
    wy = wd[:, fm:]
    r = torch.zeros((c_out, fm, 3, 5), dtype=torch.float64, device=w.device)          # index = delta + 2
    for kd in range(3):
        for kw in range(3):
            r[:, :, :, kw - kd + 2] += wy[:, :, kd, :, kw]
    wab = torch.zeros((2 * c_out, fm, 3, 3, 3), dtype=torch.float32, device=w.device)
    wab[:c_out, :, 1].copy_(r[:, :, :, 0:3].float())
    wab[c_out:, :, 1, :, 1:3].copy_(r[:, :, :, 3:5].float())
    return wl, wab


def _prod(t):
    r = 1
    for v in t:
        r *= int(v)
    return r


# --------------------------------------------------------------------------------------------------------
# module-level entry points used by the nn.Module boundary
# --------------------------------------------------------------------------------------------------------

DEFAULT_OPTIONS = {"planes": 2, "conv": "tc", "mma_terms": 0, "fuse": True, "fuse_cv": True, "feature": "native",
                   "assume_frozen": False}


# tensor-core kernel / plan-rewrite knobs an ``engine_options`` dict may carry (defaults = what the product runs)
_TC_KNOBS = {"accum_split": 0, "acc_sets": 0, "fuse_cv": True, "fuse_head": True, "collapse_stem0": True, "tile_w_log2": 0,
             "resident_weights": 0, "depth_chunk": 0, "reuse_buffers": True,
             "fuse_resample_conv": False}


def _options(model) -> dict:
    o = dict(DEFAULT_OPTIONS)
    o.update(getattr(model, "engine_options", None) or {})
    return o


def _plans(owner) -> dict:
    d = owner.__dict__.get("_lea_plans")
    if d is None:
        d = {}
        owner.__dict__["_lea_plans"] = d
    return d


def invalidate_cached_plans(model):
    for mod in model.modules():
        mod.__dict__.pop("_lea_plans", None)


def get_plan(matching: newMatching, B: int, spatial, device, options: dict, ops: Optional[Ops] = None) -> MatchingPlan:
    ops = ops or get_ops()
    knobs = {k: options.get(k, d) for k, d in _TC_KNOBS.items()}
    key = (str(device), B, tuple(spatial), options["planes"], options["conv"], options["mma_terms"],
           bool(options.get("fuse", True)), tuple(sorted(knobs.items())), id(ops))
    with _LOCK:
        plans = _plans(matching)
        plan = plans.get(key)
        if plan is None:
            plan = MatchingPlan(matching, ops, B, spatial, options["planes"], device, options["conv"],
                                options["mma_terms"], bool(options.get("fuse", True)), knobs)
            plans[key] = plan
    return plan


def _require_inference(module):
    if module.training:
        raise NotImplementedError(
            "leastereo_b200: this entry point is inference-only; train-mode BatchNorm and the backward kernels run "
            "through LEAStereo.forward (leastereo_b200/training.py) - there is deliberately no PyTorch fallback")


def _no_eval_autograd(*inputs):
    """The eval-mode path is inference-only (its result is detached).  Refuse silently wrong gradients: an input that
    requires grad under ``enable_grad`` would get none.  (Train mode - ``model.train()`` - is differentiable.)"""
    if torch.is_grad_enabled() and any(t.requires_grad for t in inputs):
        raise RuntimeError("leastereo_b200: the eval-mode forward is not differentiable (inputs require grad); wrap the "
                           "call in torch.no_grad() as predict.py / train.py's val() do, or switch the module to "
                           "train() for the differentiable path")


def hot_path_forward(model, fx: torch.Tensor, fy: torch.Tensor, ops: Optional[Ops] = None) -> torch.Tensor:
    """cost volume -> matching net -> disparity head on feature maps (B, C, H3, W3) -> (B, 3*H3, 3*W3)."""
    if model.training:
        from .training import hot_path_train_forward      # batch-statistics BN + backward kernels
        return hot_path_train_forward(model, fx, fy, ops)
    ops = ops or get_ops()
    opt = _options(model)
    fx = fx.detach().float().contiguous()
    fy = fy.detach().float().contiguous()
    B, Cn, H3, W3 = fx.shape
    D3 = int(model.maxdisp / 3)
    if D3 < 1:
        raise LeaError("maxdisp %r gives an empty cost volume" % (model.maxdisp,))
    plan = get_plan(model.matching, B, (D3, H3, W3), fx.device, opt, ops)
    if plan.fxp is not None:                 # fused: stem0's loader builds the volume from the packed feature maps
        ops.pack(fx, opt["planes"], out=plan.fxp)
        ops.pack(fy, opt["planes"], out=plan.fyp)
        if plan.fxy3 is not None:            # exact copies for the 2-D maps of the collapsed stem0
            ops.pack(fx, 3, out=PlanesVol(plan.fxy3.t[:B]))
            ops.pack(fy, 3, out=PlanesVol(plan.fxy3.t[B:]))
    else:
        ops.cost_volume_planes(fx, fy, model.maxdisp, opt["planes"], out=plan.cost)
    mat = plan.run(check_params=not opt["assume_frozen"])
    return ops.disp_head(mat, model.maxdisp)


def full_forward(model, left: torch.Tensor, right: torch.Tensor, ops: Optional[Ops] = None) -> Optional[torch.Tensor]:
    """Feature net + hot path entirely on the native kernels (eval mode).  Returns None when the configuration is not
    taken by the native feature path (the caller then runs the stock-PyTorch feature net + ``hot_path_forward``)."""
    opt = _options(model)
    if model.training or opt.get("feature", "native") != "native" or opt["conv"] != "tc" or left.shape != right.shape:
        return None
    ops = ops or get_ops()
    _no_eval_autograd(left, right)
    B, _, H, W = left.shape
    D3 = int(model.maxdisp / 3)
    h3, w3 = (H - 1) // 3 + 1, (W - 1) // 3 + 1
    if D3 < 1:
        raise LeaError("maxdisp %r gives an empty cost volume" % (model.maxdisp,))
    plan = get_plan(model.matching, B, (D3, h3, w3), left.device, opt, ops)
    if plan.fxy is None:
        return None
    fterms = int(opt.get("feature_terms", opt["mma_terms"]))
    key = ("feature", str(left.device), B, H, W, int(opt.get("feature_planes", 3)), fterms, id(ops), id(plan))
    with _LOCK:
        plans = _plans(model.feature)
        fplan = plans.get(key)
        if fplan is None:
            try:
                fplan = FeaturePlan(model.feature, ops, 2 * B, H, W, int(opt.get("feature_planes", 3)), left.device, plan.fxy,
                                    fterms, bool(opt.get("fuse", True)),
                                    {k: opt.get(k, d) for k, d in _TC_KNOBS.items()}, out3=plan.fxy3)
            except LeaError:
                fplan = False
            plans[key] = fplan
    if fplan is False:
        return None
    fplan.img[:B].copy_(left.detach().float())
    fplan.img[B:].copy_(right.detach().float())
    if not (opt.get("cuda_graph", True) and left.is_cuda) or torch.cuda.is_current_stream_capturing():
        fplan.run(check_params=not opt["assume_frozen"])
        mat = plan.run(check_params=not opt["assume_frozen"])
        return ops.disp_head(mat, model.maxdisp)
    # The module replays its own CUDA graph of the ~160 launches (predict.py-style callers get the graph's rate without
    # writing capture code: the eager launch list is host-bound at small batches - 164 against 209 pairs/s at batch 1).
    # Parameters are checked on the host before every replay; the packed weight images and BN vectors are rewritten in
    # place, and a change of the parameter key (values or storage) drops the graph so that raw weight pointers recorded
    # in it can never go stale.  Two eager calls first (allocations, lazily built tensor maps).
    st = fplan.__dict__.setdefault("_graph", {"calls": 0, "graph": None, "out": None, "keys": None, "failed": None})
    if not opt["assume_frozen"] or plan._param_key is None or fplan._param_key is None:
        fplan.refresh_params()
        plan.refresh_params()
    keys = (fplan._param_key, plan._param_key)
    if st["graph"] is not None and st["keys"] != keys:
        st["graph"], st["out"] = None, None
    if st["graph"] is None:
        st["calls"] += 1
        if st["calls"] <= 2 or st["failed"] is not None:
            fplan.run(check_params=False)
            return ops.disp_head(plan.run(check_params=False), model.maxdisp)
        torch.cuda.synchronize(left.device)
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                fplan.run(check_params=False)
                st["out"] = ops.disp_head(plan.run(check_params=False), model.maxdisp)
        except RuntimeError as e:      # capture refused by the runtime: stay on the (same) eager launch list, loudly
            import warnings
            st["failed"] = str(e)
            warnings.warn("leastereo_b200: CUDA graph capture of the forward failed (%s); launching eagerly" % str(e)[:200])
            torch.cuda.synchronize(left.device)
            fplan.run(check_params=False)
            return ops.disp_head(plan.run(check_params=False), model.maxdisp)
        st["graph"], st["keys"] = g, keys
    st["graph"].replay()
    return st["out"].clone()


def matching_forward(matching: newMatching, x: torch.Tensor, ops: Optional[Ops] = None,
                     options: Optional[dict] = None) -> torch.Tensor:
    """``newMatching.forward`` on a materialised fp32 (B, 2C, D3, H3, W3) cost volume."""
    _require_inference(matching)
    ops = ops or get_ops()
    opt = dict(DEFAULT_OPTIONS)
    opt.update(options or {})
    x = x.detach().float().contiguous()
    B, _, D, H, W = x.shape
    opt["fuse_cv"] = False                   # the caller hands over a materialised volume
    plan = get_plan(matching, B, (D, H, W), x.device, opt, ops)
    ops.pack(x, opt["planes"], out=plan.cost)
    return plan.run(check_params=not opt["assume_frozen"]).clone()


def conv_br_forward(mod: ConvBR3d, x: torch.Tensor, ops: Optional[Ops] = None, planes: int = 3) -> torch.Tensor:
    """Stand-alone ``ConvBR.forward`` (operations_3d.py:41-47) on fp32 NCDHW input, eval-mode BN."""
    _require_inference(mod)
    ops = ops or get_ops()
    x = x.detach().float().contiguous()
    w = mod.conv.weight.detach()
    c_out, c_in, k = w.shape[0], w.shape[1], w.shape[2]
    pad_in = (-c_in) % 8
    if pad_in:
        x = torch.nn.functional.pad(x, (0, 0, 0, 0, 0, 0, 0, pad_in))
        w = torch.nn.functional.pad(w, (0, 0, 0, 0, 0, 0, 0, pad_in))
    src = ops.pack(x, planes)
    out = torch.empty((x.shape[0], c_out) + tuple(x.shape[2:]), dtype=torch.float32, device=x.device)
    scale = shift = None
    if mod.use_bn:
        bn = mod.bn
        scale = (bn.weight.detach() / torch.sqrt(bn.running_var + bn.eps)).float().contiguous()
        shift = (bn.bias.detach() - bn.running_mean * scale).float().contiguous()
    p = ops.make_conv(src, 0, c_in + pad_in, c_out, k, scale, shift, mod.relu, dst_f32=out)
    ops.conv3d_simt(p, w.contiguous(), x)
    return out


def disp_head_forward(x: torch.Tensor, maxdisp: int, ops: Optional[Ops] = None) -> torch.Tensor:
    return (ops or get_ops()).disp_head(x.detach(), maxdisp)


def disparity_regression(x: torch.Tensor, maxdisp: int, ops: Optional[Ops] = None) -> torch.Tensor:
    return (ops or get_ops()).disparity_regression(x.detach(), maxdisp)
