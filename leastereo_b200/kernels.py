"""ctypes binding of the C-ABI library (``include/leastereo_b200.h``) and thin tensor-level wrappers.

PyTorch is plumbing here: it allocates device memory and supplies the current CUDA stream; all arithmetic of the
hot path happens inside ``libleastereo_b200.so``.  There is no fallback: if the library has not been built
(``python -c 'import __graft_entry__ as g; g.build()'``) or a tensor is not on a CUDA device, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_C", "libleastereo_b200.so")
ABI_VERSION = 3


class lea_vol(C.Structure):
    _fields_ = [("data", C.c_void_p), ("B", C.c_int32), ("C", C.c_int32), ("P", C.c_int32),
                ("D", C.c_int32), ("H", C.c_int32), ("W", C.c_int32)]


class lea_conv(C.Structure):
    _fields_ = [("src", lea_vol), ("src_c0", C.c_int32),
                ("c_in", C.c_int32), ("c_out", C.c_int32), ("ksize", C.c_int32),
                ("bn_scale", C.c_void_p), ("bn_shift", C.c_void_p), ("relu", C.c_int32),
                ("has_res", C.c_int32), ("res", lea_vol), ("res_c0", C.c_int32),
                ("dst", lea_vol), ("dst_c0", C.c_int32), ("dst_f32", C.c_void_p)]


class lea_rc_out(C.Structure):
    _fields_ = [("dst", lea_vol), ("dst_c0", C.c_int32), ("c_out", C.c_int32), ("weight", C.c_void_p),
                ("bn_scale", C.c_void_p), ("bn_shift", C.c_void_p), ("relu", C.c_int32)]


class lea_tc_opts(C.Structure):
    _fields_ = [("mma_terms", C.c_int32), ("fused_cv", C.c_int32), ("fx", lea_vol), ("fy", lea_vol),
                ("d3", C.c_int32), ("num_sms", C.c_int32), ("accum_split", C.c_int32), ("acc_sets", C.c_int32), ("cv_maps", C.c_void_p),
                ("resident_weights", C.c_int32), ("cv_skip", C.c_int32), ("debug", C.c_int32), ("depth_chunk", C.c_int32),
                ("tile_w_log2", C.c_int32)]


# every symbol include/leastereo_b200.h declares: name -> (restype, argtypes)
_i32, _vp, _i64 = C.c_int32, C.c_void_p, C.c_int64
_VOLP, _CONVP, _TCP = C.POINTER(lea_vol), C.POINTER(lea_conv), C.POINTER(lea_tc_opts)
SYMBOLS = {
    "lea_abi_version": (C.c_int, []),
    "lea_last_error": (C.c_char_p, []),
    "lea_is_device_build": (C.c_int, []),
    "lea_cost_volume_f32": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "lea_cost_volume_planes": (C.c_int, [_vp, _vp, _VOLP, _i32, _vp]),
    "lea_pack_planes": (C.c_int, [_vp, _VOLP, _i32, _i32, _vp]),
    "lea_unpack_planes": (C.c_int, [_VOLP, _i32, _i32, _vp, _vp]),
    "lea_trilinear_ac": (C.c_int, [_VOLP, _i32, _VOLP, _i32, _i32, _vp, _vp, _i32, _vp]),
    "lea_resample_conv1x1": (C.c_int, [_VOLP, _i32, _i32, _vp, _i32, _vp]),
    "lea_conv3d_simt": (C.c_int, [_CONVP, _vp, _vp]),
    "lea_tc_weight_image_bytes": (_i64, [_i32, _i32, _i32, _i32]),
    "lea_pack_weights_tc": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "lea_pack_weights_tc_dgrad": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "lea_conv3d_tc": (C.c_int, [_CONVP, _vp, _TCP, _vp]),
    "lea_tc_selftest": (C.c_int, [_i32, _vp]),
    "lea_fused_cv_maps_bytes": (_i64, [_i32]),
    "lea_build_fused_cv_maps": (C.c_int, [_VOLP, _VOLP, _i32, _vp, _vp]),
    "lea_stem0_assemble": (C.c_int, [_VOLP, _VOLP, _VOLP, _i32, _i32, _vp, _vp, _i32, _vp]),
    "lea_head_taps_workspace_bytes": (_i64, [_i32, _i32, _i32, _i32, _i32, _i32]),
    "lea_head_taps": (C.c_int, [_VOLP, _i32, _vp, _i32, _i32, _i32, _vp, _vp]),
    "lea_disp_head": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "lea_disparity_regression": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "lea_feature_stem": (C.c_int, [_vp, _i32, _i32, _i32, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _i32, _VOLP, _i32, _vp]),
    "lea_channel_reduce": (C.c_int, [_VOLP, _i32, _VOLP, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _i32, _vp]),
    "lea_bn_finalize": (C.c_int, [_vp, _i32, _i32, C.c_double, _vp, _vp, C.c_double, C.c_double, _vp, _vp, _vp, _vp, _vp,
                                  _vp, _vp, _vp]),
    "lea_bn_bwd_coeffs": (C.c_int, [_vp, _i32, _i32, C.c_double, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp]),
    "lea_affine_relu": (C.c_int, [_VOLP, _i32, _VOLP, _i32, _i32, _vp, _vp, _i32, _i32, _vp]),
    "lea_bn_relu_bwd": (C.c_int, [_VOLP, _i32, _VOLP, _i32, _VOLP, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "lea_conv3d_wgrad": (C.c_int, [_VOLP, _i32, _i32, _VOLP, _i32, _i32, _i32, _vp, _vp]),
    "lea_conv3d_wgrad_tc_supported": (C.c_int, [_i32, _i32, _i32, _i32]),
    "lea_conv3d_wgrad_tc": (C.c_int, [_VOLP, _i32, _i32, _VOLP, _i32, _i32, _i32, _vp, _vp]),
    "lea_trilinear_ac_bwd": (C.c_int, [_VOLP, _i32, _VOLP, _i32, _i32, _i32, _vp]),
    "lea_cost_volume_bwd": (C.c_int, [_VOLP, _i32, _vp, _vp, _vp]),
    "lea_disp_head_bwd": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "lea_image_stats_u8": (C.c_int, [_vp, _i32, _i32, _vp, _vp]),
    "lea_normalize_pad_u8": (C.c_int, [_vp, _i32, _i32, _vp, _vp, _i32, _i32, _vp]),
    "lea_masked_smooth_l1": (C.c_int, [_vp, _vp, _i64, C.c_float, _vp, _vp]),
    "lea_masked_smooth_l1_bwd": (C.c_int, [_vp, _vp, _i64, C.c_float, _vp, C.c_float, _vp, _vp]),
    "lea_adam_step": (C.c_int, [_vp, _vp, _vp, _vp, _i64, C.c_float, C.c_float, C.c_float, C.c_float, _i32, _vp]),
    "lea_disparity_metrics": (C.c_int, [_vp, _vp, _i64, C.c_float, _vp, C.c_int32, _vp, _vp]),
}
# symbols only the CUDA build has (tcgen05 path); the CPU emulation used by the no-GPU tests lacks them
DEVICE_ONLY = {"lea_tc_weight_image_bytes", "lea_pack_weights_tc", "lea_pack_weights_tc_dgrad", "lea_conv3d_tc", "lea_tc_selftest",
               "lea_fused_cv_maps_bytes", "lea_build_fused_cv_maps", "lea_conv3d_wgrad_tc_supported", "lea_conv3d_wgrad_tc"}


class LeaError(RuntimeError):
    pass


class PlanesVol:
    """Python handle of a planes volume: bf16 tensor of shape (B, C/8, P, D, H, W, 8)."""

    __slots__ = ("t", "B", "C", "P", "D", "H", "W", "_arena_index")

    def __init__(self, t: torch.Tensor):
        assert t.dtype == torch.bfloat16 and t.dim() == 7 and t.shape[-1] == 8 and t.is_contiguous()
        self.t = t
        self._arena_index = None        # set by engine.MatchingPlan for volumes that live in its reuse arena
        self.B, cb, self.P, self.D, self.H, self.W, _ = t.shape
        self.C = cb * 8

    @staticmethod
    def empty(B, Cn, P, D, H, W, device) -> "PlanesVol":
        assert Cn % 8 == 0
        return PlanesVol(torch.empty((B, Cn // 8, P, D, H, W, 8), dtype=torch.bfloat16, device=device))

    @property
    def spatial(self):
        return (self.D, self.H, self.W)

    def struct(self) -> lea_vol:
        ptr = 0 if self.t.is_meta else self.t.data_ptr()        # meta: the engine's dry-run build (sizes only)
        return lea_vol(ptr, self.B, self.C, self.P, self.D, self.H, self.W)

    def nbytes(self) -> int:
        return self.t.numel() * 2


class Ops:
    """All entry points of one loaded library."""

    def __init__(self, lib_path: str = LIB_PATH, require_device_build: bool = True):
        if not os.path.exists(lib_path):
            raise LeaError(
                "leastereo_b200: native library %s is missing - build it with "
                "`python -c \"import __graft_entry__ as g; g.build()\"`; there is no PyTorch/CPU fallback for the "
                "hot path" % lib_path)
        self.lib = C.CDLL(lib_path)
        self.lib_path = lib_path
        missing = []
        for name, (res, args) in SYMBOLS.items():
            try:
                fn = getattr(self.lib, name)
            except AttributeError:
                missing.append(name)
                continue
            fn.restype = res
            fn.argtypes = args
        self.missing = missing
        if self.lib.lea_abi_version() != ABI_VERSION:
            raise LeaError("ABI version mismatch: library %d, binding %d" % (self.lib.lea_abi_version(), ABI_VERSION))
        self.device_build = bool(self.lib.lea_is_device_build())
        if require_device_build:
            if not self.device_build:
                raise LeaError("%s is not a device build" % lib_path)
            if missing:
                raise LeaError("%s lacks symbols %r" % (lib_path, missing))
        self.launches = 0          # kernels launched through this binding (bench.py reports it)

    # ---- helpers ----------------------------------------------------------------------------------------
    def _check(self, rc: int):
        self.launches += 1
        if rc != 0:
            raise LeaError(self.lib.lea_last_error().decode())

    def _dev(self, *tensors):
        for t in tensors:
            if self.device_build and not t.is_cuda:
                raise LeaError("leastereo_b200 kernels need CUDA tensors (got %s); no CPU fallback exists" % t.device)
            if not self.device_build and t.is_cuda:
                raise LeaError("emulation library used with CUDA tensors")

    def _stream(self, t: torch.Tensor):
        if t.is_cuda:
            return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)
        return C.c_void_p(0)

    @staticmethod
    def _f32(t: torch.Tensor) -> torch.Tensor:
        if t.dtype != torch.float32:
            raise LeaError("expected float32, got %s" % t.dtype)
        return t.contiguous()

    # ---- cost volume ------------------------------------------------------------------------------------
    def cost_volume_f32(self, x: torch.Tensor, y: torch.Tensor, maxdisp: int) -> torch.Tensor:
        x, y = self._f32(x), self._f32(y)
        self._dev(x, y)
        assert x.shape == y.shape and x.dim() == 4
        B, Cn, H, W = x.shape
        D3 = int(maxdisp / 3)
        cost = torch.empty((B, 2 * Cn, D3, H, W), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device) if x.is_cuda else _null():
            self._check(self.lib.lea_cost_volume_f32(x.data_ptr(), y.data_ptr(), cost.data_ptr(), B, Cn, H, W, D3,
                                                     self._stream(x)))
        return cost

    def cost_volume_planes(self, x: torch.Tensor, y: torch.Tensor, maxdisp: int, P: int,
                           out: Optional[PlanesVol] = None) -> PlanesVol:
        x, y = self._f32(x), self._f32(y)
        self._dev(x, y)
        B, Cn, H, W = x.shape
        D3 = int(maxdisp / 3)
        vol = out or PlanesVol.empty(B, 2 * Cn, P, D3, H, W, x.device)
        vs = vol.struct()
        with torch.cuda.device(x.device) if x.is_cuda else _null():
            self._check(self.lib.lea_cost_volume_planes(x.data_ptr(), y.data_ptr(), C.byref(vs), Cn, self._stream(x)))
        return vol

    # ---- planes pack / unpack ---------------------------------------------------------------------------
    def pack(self, src: torch.Tensor, P: int, out: Optional[PlanesVol] = None, c0: int = 0) -> PlanesVol:
        src = self._f32(src)
        self._dev(src)
        if src.dim() == 4:
            src = src.unsqueeze(2)
        B, Cn, D, H, W = src.shape
        vol = out or PlanesVol.empty(B, Cn, P, D, H, W, src.device)
        vs = vol.struct()
        with torch.cuda.device(src.device) if src.is_cuda else _null():
            self._check(self.lib.lea_pack_planes(src.data_ptr(), C.byref(vs), c0, Cn, self._stream(src)))
        return vol

    def unpack(self, vol: PlanesVol, c0: int = 0, c: Optional[int] = None) -> torch.Tensor:
        c = vol.C - c0 if c is None else c
        out = torch.empty((vol.B, c, vol.D, vol.H, vol.W), dtype=torch.float32, device=vol.t.device)
        vs = vol.struct()
        with torch.cuda.device(out.device) if out.is_cuda else _null():
            self._check(self.lib.lea_unpack_planes(C.byref(vs), c0, c, out.data_ptr(), self._stream(out)))
        return out

    # ---- trilinear --------------------------------------------------------------------------------------
    def trilinear_ac(self, src: PlanesVol, src_c0: int, c: int, dst: PlanesVol, dst_c0: int = 0,
                     bn_scale: Optional[torch.Tensor] = None, bn_shift: Optional[torch.Tensor] = None,
                     relu: bool = False):
        self._dev(src.t, dst.t)
        s, d = src.struct(), dst.struct()
        sc = bn_scale.data_ptr() if bn_scale is not None else None
        sh = bn_shift.data_ptr() if bn_shift is not None else None
        with torch.cuda.device(src.t.device) if src.t.is_cuda else _null():
            self._check(self.lib.lea_trilinear_ac(C.byref(s), src_c0, C.byref(d), dst_c0, c, sc, sh, int(bool(relu)),
                                                  self._stream(src.t)))

    def resample_conv1x1(self, src: PlanesVol, src_c0: int, c_in: int, outs):
        """outs: list of (dst PlanesVol, dst_c0, c_out, weight fp32 (c_out, c_in, ...), bn_scale ptr|tensor|None,
        bn_shift, relu) - one or two 1x1x1 ConvBR consumers of the align_corners=True resample of src."""
        self._dev(src.t)
        arr = (lea_rc_out * len(outs))()
        for k, (dst, dst_c0, c_out, weight, sc, sh, relu) in enumerate(outs):
            assert weight.dtype == torch.float32 and weight.is_contiguous() and weight.numel() == c_out * c_in
            arr[k].dst, arr[k].dst_c0, arr[k].c_out = dst.struct(), dst_c0, c_out
            arr[k].weight = weight.data_ptr()
            arr[k].bn_scale = sc if isinstance(sc, int) or sc is None else sc.data_ptr()
            arr[k].bn_shift = sh if isinstance(sh, int) or sh is None else sh.data_ptr()
            arr[k].relu = int(bool(relu))
        s = src.struct()
        with self._dev_ctx(src.t):
            self._check(self.lib.lea_resample_conv1x1(C.byref(s), src_c0, c_in, arr, len(outs), self._stream(src.t)))

    # ---- conv -------------------------------------------------------------------------------------------
    def make_conv(self, src: PlanesVol, src_c0: int, c_in: int, c_out: int, ksize: int,
                  bn_scale: Optional[torch.Tensor], bn_shift: Optional[torch.Tensor], relu: bool,
                  dst: Optional[PlanesVol] = None, dst_c0: int = 0, res: Optional[PlanesVol] = None, res_c0: int = 0,
                  dst_f32: Optional[torch.Tensor] = None) -> lea_conv:
        p = lea_conv()
        p.src, p.src_c0, p.c_in, p.c_out, p.ksize = src.struct(), src_c0, c_in, c_out, ksize
        p.bn_scale = bn_scale.data_ptr() if bn_scale is not None else None
        p.bn_shift = bn_shift.data_ptr() if bn_shift is not None else None
        p.relu = int(bool(relu))
        p.has_res = int(res is not None)
        if res is not None:
            p.res, p.res_c0 = res.struct(), res_c0
        if dst_f32 is not None:
            assert dst_f32.dtype == torch.float32 and dst_f32.is_contiguous()
            p.dst_f32 = dst_f32.data_ptr()
        else:
            p.dst, p.dst_c0 = dst.struct(), dst_c0
            p.dst_f32 = None
        return p

    def conv3d_simt(self, p: lea_conv, weight: torch.Tensor, ref: torch.Tensor):
        self._dev(weight, ref)
        assert weight.dtype == torch.float32 and weight.is_contiguous()
        with torch.cuda.device(ref.device) if ref.is_cuda else _null():
            self._check(self.lib.lea_conv3d_simt(C.byref(p), weight.data_ptr(), self._stream(ref)))

    def tc_weight_image_bytes(self, c_in, c_out, ksize, planes) -> int:
        return int(self.lib.lea_tc_weight_image_bytes(c_in, c_out, ksize, planes))

    def pack_weights_tc(self, weight: torch.Tensor, planes: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        self._dev(weight)
        weight = self._f32(weight)
        c_out, c_in, k = weight.shape[0], weight.shape[1], weight.shape[2]
        nbytes = self.tc_weight_image_bytes(c_in, c_out, k, planes)
        if nbytes <= 0:
            raise LeaError("tcgen05 conv does not take c_in=%d c_out=%d k=%d" % (c_in, c_out, k))
        img = out if out is not None else torch.empty(nbytes, dtype=torch.uint8, device=weight.device)
        assert img.numel() == nbytes
        with torch.cuda.device(weight.device):
            self._check(self.lib.lea_pack_weights_tc(weight.data_ptr(), img.data_ptr(), c_in, c_out, k, planes,
                                                     self._stream(weight)))
        return img

    def pack_weights_tc_dgrad(self, weight: torch.Tensor, planes: int, ci0: int = 0, ci: Optional[int] = None,
                              out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Image of the data-gradient conv (channels transposed, taps flipped) straight from the forward weight
        (c_out_fwd, c_in_fwd, k, k, k); ``[ci0, ci0 + ci)`` = the forward input channels this launch produces."""
        self._dev(weight)
        assert weight.dtype == torch.float32 and weight.is_contiguous()
        fwd_c_in = weight.shape[1]
        c_in, k = weight.shape[0], weight.shape[2]                               # of the data-gradient conv
        c_out = fwd_c_in - ci0 if ci is None else ci
        nbytes = self.tc_weight_image_bytes(c_in, c_out, k, planes)
        if nbytes <= 0:
            raise LeaError("tcgen05 conv does not take c_in=%d c_out=%d k=%d" % (c_in, c_out, k))
        img = out if out is not None else torch.empty(nbytes, dtype=torch.uint8, device=weight.device)
        assert img.numel() == nbytes
        with torch.cuda.device(weight.device):
            self._check(self.lib.lea_pack_weights_tc_dgrad(weight.data_ptr(), img.data_ptr(), c_in, c_out, k, planes,
                                                           fwd_c_in, ci0, self._stream(weight)))
        return img

    def conv3d_tc(self, p: lea_conv, wimg: torch.Tensor, opts: lea_tc_opts, ref: torch.Tensor):
        self._dev(wimg, ref)
        with torch.cuda.device(ref.device):
            self._check(self.lib.lea_conv3d_tc(C.byref(p), wimg.data_ptr(), C.byref(opts), self._stream(ref)))

    def build_fused_cv_maps(self, fx: PlanesVol, fy: PlanesVol, d3: int) -> torch.Tensor:
        """Device array of per-disparity TMA descriptors for the fused cost-volume loader of stem0."""
        self._dev(fx.t, fy.t)
        nbytes = int(self.lib.lea_fused_cv_maps_bytes(d3))
        maps = torch.empty(nbytes + 128, dtype=torch.uint8, device=fx.t.device)
        off = (-maps.data_ptr()) % 128                      # descriptors must be 64-byte aligned; use 128
        maps = maps[off: off + nbytes]
        a, b = fx.struct(), fy.struct()
        with torch.cuda.device(fx.t.device):
            self._check(self.lib.lea_build_fused_cv_maps(C.byref(a), C.byref(b), d3, maps.data_ptr(),
                                                         self._stream(fx.t)))
        return maps

    def tc_selftest(self, verbose: int = 1) -> int:
        return int(self.lib.lea_tc_selftest(verbose, C.c_void_p(torch.cuda.current_stream().cuda_stream)))

    # ---- collapsed stem0 ---------------------------------------------------------------------------------------
    def stem0_assemble(self, lmap: PlanesVol, abmap: PlanesVol, dst: PlanesVol, dst_c0: int, c_out: int,
                       bn_scale: Optional[torch.Tensor], bn_shift: Optional[torch.Tensor], relu: bool):
        self._dev(lmap.t, abmap.t, dst.t)
        a, b, d = lmap.struct(), abmap.struct(), dst.struct()
        with self._dev_ctx(dst.t):
            self._check(self.lib.lea_stem0_assemble(C.byref(a), C.byref(b), C.byref(d), dst_c0, c_out,
                                                    self._ptr(bn_scale), self._ptr(bn_shift), int(bool(relu)),
                                                    self._stream(dst.t)))

    # ---- matching-net head (upsample_6 -> last_3 without the up-sampled volume) ----------------------------
    def head_taps_workspace(self, q: PlanesVol, spatial) -> torch.Tensor:
        D, H, W = (int(v) for v in spatial)
        n = int(self.lib.lea_head_taps_workspace_bytes(q.B, q.D, q.H, D, H, W))
        return torch.empty(max(n // 4, 1), dtype=torch.float32, device=q.t.device)

    def head_taps(self, q: PlanesVol, q_c0: int, mat: torch.Tensor, workspace: Optional[torch.Tensor] = None):
        """mat (B, 1, D, H, W) fp32 <- sum over the 27 tap channels of q, up-sampled and shifted (see the header)."""
        self._dev(q.t, mat)
        assert mat.dtype == torch.float32 and mat.is_contiguous() and mat.dim() == 5 and mat.shape[1] == 1
        D, H, W = (int(v) for v in mat.shape[2:])
        ws = workspace if workspace is not None else self.head_taps_workspace(q, (D, H, W))
        qs = q.struct()
        with torch.cuda.device(mat.device) if mat.is_cuda else _null():
            self._check(self.lib.lea_head_taps(C.byref(qs), q_c0, mat.data_ptr(), D, H, W, ws.data_ptr(),
                                               self._stream(mat)))
        self.launches += 2                      # three kernels behind one entry point

    # ---- disparity head ---------------------------------------------------------------------------------
    def disp_head(self, mat: torch.Tensor, maxdisp: int) -> torch.Tensor:
        mat = self._f32(mat)
        self._dev(mat)
        if mat.dim() == 5:
            assert mat.shape[1] == 1
            mat = mat[:, 0]
        B, D3, H3, W3 = mat.shape
        out = torch.empty((B, 3 * H3, 3 * W3), dtype=torch.float32, device=mat.device)
        with torch.cuda.device(mat.device) if mat.is_cuda else _null():
            self._check(self.lib.lea_disp_head(mat.data_ptr(), out.data_ptr(), B, D3, H3, W3, int(maxdisp),
                                               self._stream(mat)))
        return out

    def feature_stem(self, img: torch.Tensor, w0, sc0, sh0, w1, sc1, sh1, dst: PlanesVol, dst_c0: int = 0):
        img = self._f32(img)
        self._dev(img, dst.t)
        B, _, H, W = img.shape
        d = dst.struct()
        with torch.cuda.device(img.device) if img.is_cuda else _null():
            self._check(self.lib.lea_feature_stem(img.data_ptr(), B, H, W, w0.data_ptr(), sc0.data_ptr(), sh0.data_ptr(),
                                                  w0.shape[0], w1.data_ptr(), sc1.data_ptr(), sh1.data_ptr(), w1.shape[0],
                                                  C.byref(d), dst_c0, self._stream(img)))

    # ---- training side ------------------------------------------------------------------------------------
    def _dev_ctx(self, t: torch.Tensor):
        return torch.cuda.device(t.device) if t.is_cuda else _null()

    @staticmethod
    def _ptr(t: Optional[torch.Tensor]):
        return None if t is None else t.data_ptr()

    def channel_reduce(self, x: PlanesVol, x_c0: int, c: int, mode: int = 0, dy: Optional[PlanesVol] = None,
                       dy_c0: int = 0, relu: bool = False, scale=None, shift=None, mean=None, invstd=None,
                       chunks: int = 0) -> torch.Tensor:
        """Per-channel (sum, sum of squares) [mode 0] or BN-backward sums [mode 1]; returns float64 (2, c)."""
        self._dev(x.t)
        vox = x.B * x.D * x.H * x.W
        if chunks <= 0:
            chunks = max(64, 1184 // max(1, c >> 3))          # ~8 CTAs per SM over all channel blocks
        chunks = max(1, min(chunks, (vox // max(1, x.B) + 255) // 256))
        partial = torch.empty((chunks, 2, c), dtype=torch.float32, device=x.t.device)
        xs = x.struct()
        ds = dy.struct() if dy is not None else xs
        with self._dev_ctx(x.t):
            self._check(self.lib.lea_channel_reduce(C.byref(xs), x_c0, C.byref(ds), dy_c0, c, mode, int(bool(relu)),
                                                    self._ptr(scale), self._ptr(shift), self._ptr(mean),
                                                    self._ptr(invstd), partial.data_ptr(), chunks, self._stream(x.t)))
        return partial.double().sum(dim=0)

    def channel_reduce_partial(self, x: PlanesVol, x_c0: int, c: int, partial: torch.Tensor, mode: int = 0,
                               dy: Optional[PlanesVol] = None, dy_c0: int = 0, relu: bool = False, scale=None,
                               shift=None, mean=None, invstd=None) -> int:
        """``lea_channel_reduce`` into the caller's (chunks, 2, c) fp32 buffer; returns the number of chunk rows written
        (the rows are added on the device by ``bn_finalize`` / ``bn_bwd_coeffs``)."""
        self._dev(x.t, partial)
        chunks = max(64, 1184 // max(1, c >> 3))
        chunks = max(1, min(chunks, ((x.D * x.H * x.W) + 255) // 256, partial.shape[0]))
        assert partial.dtype == torch.float32 and partial.is_contiguous() and partial.shape[1:] == (2, c)
        xs = x.struct()
        ds = dy.struct() if dy is not None else xs
        with self._dev_ctx(x.t):
            self._check(self.lib.lea_channel_reduce(C.byref(xs), x_c0, C.byref(ds), dy_c0, c, mode, int(bool(relu)),
                                                    self._ptr(scale), self._ptr(shift), self._ptr(mean),
                                                    self._ptr(invstd), partial.data_ptr(), chunks, self._stream(x.t)))
        return chunks

    def bn_finalize(self, partial: torch.Tensor, chunks: int, c: int, n: float, gamma, beta, eps: float, momentum: float,
                    running_mean, running_var, num_batches_tracked, mean, invstd, scale, shift):
        self._dev(partial, mean, invstd, scale, shift)
        for t in (gamma, beta, running_mean, running_var, mean, invstd, scale, shift):
            assert t is None or (t.dtype == torch.float32 and t.is_contiguous() and t.numel() >= c)
        assert num_batches_tracked is None or num_batches_tracked.dtype == torch.int64
        with self._dev_ctx(partial):
            self._check(self.lib.lea_bn_finalize(partial.data_ptr(), chunks, c, float(n), self._ptr(gamma), self._ptr(beta),
                                                 float(eps), float(momentum), self._ptr(running_mean),
                                                 self._ptr(running_var), self._ptr(num_batches_tracked), mean.data_ptr(),
                                                 invstd.data_ptr(), scale.data_ptr(), shift.data_ptr(),
                                                 self._stream(partial)))

    def bn_bwd_coeffs(self, partial: torch.Tensor, chunks: int, c: int, n: float, gamma, invstd, ka, kb, kc, dgamma, dbeta,
                      accumulate: bool = False):
        self._dev(partial, invstd, ka, kb, kc)
        for t in (gamma, invstd, ka, kb, kc, dgamma, dbeta):
            assert t is None or (t.dtype == torch.float32 and t.is_contiguous() and t.numel() >= c)
        with self._dev_ctx(partial):
            self._check(self.lib.lea_bn_bwd_coeffs(partial.data_ptr(), chunks, c, float(n), self._ptr(gamma),
                                                   invstd.data_ptr(), ka.data_ptr(), kb.data_ptr(), kc.data_ptr(),
                                                   self._ptr(dgamma), self._ptr(dbeta), int(bool(accumulate)),
                                                   self._stream(partial)))

    def affine_relu(self, x: PlanesVol, x_c0: int, dst: PlanesVol, dst_c0: int, c: int, scale=None, shift=None,
                    relu: bool = False, accumulate: bool = False):
        self._dev(x.t, dst.t)
        xs, ds = x.struct(), dst.struct()
        with self._dev_ctx(x.t):
            self._check(self.lib.lea_affine_relu(C.byref(xs), x_c0, C.byref(ds), dst_c0, c, self._ptr(scale),
                                                 self._ptr(shift), int(bool(relu)), int(bool(accumulate)),
                                                 self._stream(x.t)))

    def bn_relu_bwd(self, x: PlanesVol, x_c0: int, dy: PlanesVol, dy_c0: int, dx: PlanesVol, dx_c0: int, c: int,
                    relu: bool, scale, shift, mean, invstd, ka, kb, kc):
        self._dev(x.t, dy.t, dx.t)
        xs, ys, ds = x.struct(), dy.struct(), dx.struct()
        with self._dev_ctx(x.t):
            self._check(self.lib.lea_bn_relu_bwd(C.byref(xs), x_c0, C.byref(ys), dy_c0, C.byref(ds), dx_c0, c,
                                                 int(bool(relu)), scale.data_ptr(), shift.data_ptr(), mean.data_ptr(),
                                                 invstd.data_ptr(), ka.data_ptr(), kb.data_ptr(), kc.data_ptr(),
                                                 self._stream(x.t)))

    def conv3d_wgrad(self, src: PlanesVol, src_c0: int, c_in: int, dout: PlanesVol, dout_c0: int, c_out: int,
                     ksize: int, dw: torch.Tensor, tensor_cores: bool = False):
        """dw += weight gradient.  ``tensor_cores``: take ``lea_conv3d_wgrad_tc`` (mma.sync on the bf16 planes) for the
        shapes it supports (device build only); everything else runs the fp32 FMA kernel."""
        self._dev(src.t, dout.t, dw)
        assert dw.dtype == torch.float32 and dw.is_contiguous() and dw.numel() == c_out * c_in * ksize ** 3
        a, b = src.struct(), dout.struct()
        with self._dev_ctx(src.t):
            if (tensor_cores and self.device_build and src.P == 2 and dout.P == 2 and
                    self.lib.lea_conv3d_wgrad_tc_supported(c_in, c_out, ksize, src.P)):
                self._check(self.lib.lea_conv3d_wgrad_tc(C.byref(a), src_c0, c_in, C.byref(b), dout_c0, c_out, ksize,
                                                         dw.data_ptr(), self._stream(src.t)))
                return
            self._check(self.lib.lea_conv3d_wgrad(C.byref(a), src_c0, c_in, C.byref(b), dout_c0, c_out, ksize,
                                                  dw.data_ptr(), self._stream(src.t)))

    def trilinear_ac_bwd(self, ddst: PlanesVol, ddst_c0: int, dsrc: PlanesVol, dsrc_c0: int, c: int,
                         accumulate: bool = True):
        self._dev(ddst.t, dsrc.t)
        a, b = ddst.struct(), dsrc.struct()
        with self._dev_ctx(ddst.t):
            self._check(self.lib.lea_trilinear_ac_bwd(C.byref(a), ddst_c0, C.byref(b), dsrc_c0, c,
                                                      int(bool(accumulate)), self._stream(ddst.t)))

    def cost_volume_bwd(self, dcost: PlanesVol, Cn: int):
        self._dev(dcost.t)
        dx = torch.empty((dcost.B, Cn, dcost.H, dcost.W), dtype=torch.float32, device=dcost.t.device)
        dy = torch.empty_like(dx)
        a = dcost.struct()
        with self._dev_ctx(dcost.t):
            self._check(self.lib.lea_cost_volume_bwd(C.byref(a), Cn, dx.data_ptr(), dy.data_ptr(), self._stream(dcost.t)))
        return dx, dy

    def disp_head_bwd(self, mat: torch.Tensor, gout: torch.Tensor, maxdisp: int) -> torch.Tensor:
        mat, gout = self._f32(mat), self._f32(gout)
        self._dev(mat, gout)
        if mat.dim() == 5:
            mat = mat[:, 0]
        B, D3, H3, W3 = mat.shape
        assert tuple(gout.shape) == (B, 3 * H3, 3 * W3)
        dmat = torch.zeros_like(mat)
        with self._dev_ctx(mat):
            self._check(self.lib.lea_disp_head_bwd(mat.data_ptr(), gout.data_ptr(), dmat.data_ptr(), B, D3, H3, W3,
                                                   int(maxdisp), self._stream(mat)))
        return dmat

    # ---- callers on either side of the path ----------------------------------------------
    def normalize_pad_u8(self, img_hwc: torch.Tensor, crop_h: int, crop_w: int, out: Optional[torch.Tensor] = None,
                         sums: Optional[torch.Tensor] = None) -> torch.Tensor:
        """uint8 (H, W, 3) image -> z-normalised, padded / centre-cropped fp32 (3, crop_h, crop_w)
        (predict.py:144-184)."""
        self._dev(img_hwc)
        assert img_hwc.dtype == torch.uint8 and img_hwc.dim() == 3 and img_hwc.shape[2] == 3 and img_hwc.is_contiguous()
        H, W = int(img_hwc.shape[0]), int(img_hwc.shape[1])
        if sums is None:
            sums = torch.zeros(6, dtype=torch.int64, device=img_hwc.device)
        else:
            sums.zero_()
        if out is None:
            out = torch.empty((3, crop_h, crop_w), dtype=torch.float32, device=img_hwc.device)
        assert out.dtype == torch.float32 and out.is_contiguous() and tuple(out.shape) == (3, crop_h, crop_w)
        with self._dev_ctx(img_hwc):
            st = self._stream(img_hwc)
            self._check(self.lib.lea_image_stats_u8(img_hwc.data_ptr(), H, W, sums.data_ptr(), st))
            self._check(self.lib.lea_normalize_pad_u8(img_hwc.data_ptr(), H, W, sums.data_ptr(), out.data_ptr(),
                                                      crop_h, crop_w, st))
        return out

    def masked_smooth_l1(self, disp: torch.Tensor, target: torch.Tensor, maxdisp: float) -> torch.Tensor:
        """float64 [sum loss, sum |d|, #valid] over mask = 0.001 < target < maxdisp (train.py:116-118,157,162)."""
        disp, target = self._f32(disp), self._f32(target)
        self._dev(disp, target)
        assert disp.shape == target.shape
        acc = torch.zeros(3, dtype=torch.float64, device=disp.device)
        with self._dev_ctx(disp):
            self._check(self.lib.lea_masked_smooth_l1(disp.data_ptr(), target.data_ptr(), disp.numel(), float(maxdisp),
                                                      acc.data_ptr(), self._stream(disp)))
        return acc

    def masked_smooth_l1_bwd(self, disp: torch.Tensor, target: torch.Tensor, maxdisp: float, acc: torch.Tensor,
                             upstream: float = 1.0) -> torch.Tensor:
        disp, target = self._f32(disp), self._f32(target)
        self._dev(disp, target, acc)
        grad = torch.empty_like(disp)
        with self._dev_ctx(disp):
            self._check(self.lib.lea_masked_smooth_l1_bwd(disp.data_ptr(), target.data_ptr(), disp.numel(),
                                                          float(maxdisp), acc.data_ptr(), float(upstream),
                                                          grad.data_ptr(), self._stream(disp)))
        return grad

    def adam_step(self, param: torch.Tensor, grad: torch.Tensor, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor,
                  lr: float, beta1: float, beta2: float, eps: float, step: int):
        self._dev(param, grad, exp_avg, exp_avg_sq)
        for t in (param, grad, exp_avg, exp_avg_sq):
            assert t.dtype == torch.float32 and t.is_contiguous() and t.numel() == param.numel()
        with self._dev_ctx(param):
            self._check(self.lib.lea_adam_step(param.data_ptr(), grad.data_ptr(), exp_avg.data_ptr(),
                                               exp_avg_sq.data_ptr(), param.numel(), float(lr), float(beta1),
                                               float(beta2), float(eps), int(step), self._stream(param)))

    def disparity_metrics(self, pred: torch.Tensor, target: torch.Tensor, maxdisp: float,
                          thresholds=(1.0, 2.0, 3.0, 5.0), float_diff: bool = False) -> torch.Tensor:
        """float64 [#valid, sum |d| (inclusive mask), #3-px-correct, #(e <= thr_k) x4, #inclusive-valid, sum |d| (strict
        mask)] - ``lea_disparity_metrics`` (utils/metrics.py:6-46, evaluation.py:290-291, train.py:203); device tensor.
        ``float_diff`` switches OFF the reference's int64 truncation of the error (not the reference's numbers)."""
        pred, target = self._f32(pred), self._f32(target)
        self._dev(pred, target)
        assert pred.shape == target.shape and len(thresholds) == 4
        acc = torch.zeros(9, dtype=torch.float64, device=pred.device)
        thr = (C.c_float * 4)(*[float(t) for t in thresholds])
        with self._dev_ctx(pred):
            self._check(self.lib.lea_disparity_metrics(pred.data_ptr(), target.data_ptr(), pred.numel(), float(maxdisp),
                                                       thr, 1 if float_diff else 0, acc.data_ptr(), self._stream(pred)))
        return acc

    def disparity_regression(self, p: torch.Tensor, maxdisp: int) -> torch.Tensor:
        p = self._f32(p)
        self._dev(p)
        B, D, H, W = p.shape
        assert D == maxdisp
        out = torch.empty((B, H, W), dtype=torch.float32, device=p.device)
        with torch.cuda.device(p.device) if p.is_cuda else _null():
            self._check(self.lib.lea_disparity_regression(p.data_ptr(), out.data_ptr(), B, D, H, W, self._stream(p)))
        return out


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


_OPS: Optional[Ops] = None


def get_ops() -> Ops:
    """The product binding: the CUDA build of the library, or an exception - never anything else."""
    global _OPS
    if _OPS is None:
        _OPS = Ops(LIB_PATH, require_device_build=True)
    return _OPS
