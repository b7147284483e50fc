"""Small, fixed launch sequence for `ncu --set full` captures: one launch each of the headline kernels at KITTI
shapes (run plain first; see B200_PROFILING.md)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from leastereo_b200.kernels import get_ops, PlanesVol, lea_tc_opts  # noqa: E402


def main():
    which = sys.argv[1:] or ["stem0", "l1res", "l1batched", "conv1", "l0res", "cv", "disp", "resample", "headtaps", "pp64",
                             "assemble", "wgrad"]
    ops = get_ops()
    dev = torch.device("cuda:0")
    B, C, H3, W3, D3, maxdisp = 1, 32, 128, 416, 64, 192

    def conv(ci, co, k, sp, res):
        src = PlanesVol.empty(B, ci, 2, *sp, dev)
        src.t.copy_(torch.randn(src.t.shape, device=dev).bfloat16() * 0.1)
        w = torch.randn(co, ci, k, k, k, device=dev) * 0.05
        scale = torch.ones(co, device=dev); shift = torch.zeros(co, device=dev)
        dst = PlanesVol.empty(B, co, 2, *sp, dev)
        dst.t.zero_()
        p = ops.make_conv(src, 0, ci, co, k, scale, shift, True, dst=dst, res=dst if res else None)
        img = ops.pack_weights_tc(w, 2)
        opts = lea_tc_opts()
        for _ in range(2):
            ops.conv3d_tc(p, img, opts, w)
        torch.cuda.synchronize()

    if "stem0" in which:
        conv(64, 32, 3, (64, 128, 416), False)
    if "conv1" in which:
        conv(128, 64, 3, (32, 64, 208), False)
    if "l1res" in which:
        conv(16, 16, 3, (32, 64, 208), True)
    if "l1batched" in which:
        conv(16, 48, 3, (32, 64, 208), False)
    if "l0res" in which:
        conv(8, 8, 3, (64, 128, 416), True)
    if "pp64" in which:
        conv(64, 16, 1, (32, 64, 208), False)
    if "wgrad" in which:
        # weight gradients at the training shapes (288x576, one sample): stem1-like 32 -> 32 at level 0, a level-1 cell op
        # 16 -> 16, and the 1x1x1 pre-processing conv 64 -> 16
        for (ci, co, k, sp) in [(32, 32, 3, (64, 96, 192)), (16, 16, 3, (32, 48, 96)), (64, 16, 1, (32, 48, 96))]:
            x = PlanesVol.empty(B, ci, 2, *sp, dev); x.t.copy_(torch.randn(x.t.shape, device=dev).bfloat16() * 0.1)
            dy = PlanesVol.empty(B, co, 2, *sp, dev); dy.t.copy_(torch.randn(dy.t.shape, device=dev).bfloat16() * 0.1)
            dw = torch.zeros(co, ci, k, k, k, device=dev)
            for _ in range(2):
                ops.conv3d_wgrad(x, 0, ci, dy, 0, co, k, dw, tensor_cores=True)
            del x, dy
    if "cv" in which:
        x = torch.randn(B, C, H3, W3, device=dev); y = torch.randn(B, C, H3, W3, device=dev)
        for _ in range(2):
            ops.cost_volume_f32(x, y, maxdisp)
    if "disp" in which:
        mat = torch.randn(B, D3, H3, W3, device=dev) * 3
        for _ in range(2):
            ops.disp_head(mat, maxdisp)
    if "assemble" in which:
        lmap = PlanesVol.empty(B, 32, 3, 1, H3, W3, dev); lmap.t.copy_(torch.randn(lmap.t.shape, device=dev).bfloat16())
        abmap = PlanesVol.empty(B, 64, 3, 1, H3, W3, dev); abmap.t.copy_(torch.randn(abmap.t.shape, device=dev).bfloat16())
        dst = PlanesVol.empty(B, 32, 2, D3, H3, W3, dev)
        sc = torch.ones(32, device=dev); sh = torch.zeros(32, device=dev)
        for _ in range(2):
            ops.stem0_assemble(lmap, abmap, dst, 0, 32, sc, sh, True)
        del lmap, abmap, dst
    if "headtaps" in which:
        q = PlanesVol.empty(B, 32, 2, 32, 64, 208, dev); q.t.copy_(torch.randn(q.t.shape, device=dev).bfloat16())
        mat = torch.empty((B, 1, D3, H3, W3), device=dev)
        ws = ops.head_taps_workspace(q, (D3, H3, W3))
        for _ in range(2):
            ops.head_taps(q, 0, mat, ws)
    if "resample" in which:
        src = PlanesVol.empty(B, 32, 2, 64, 128, 416, dev); src.t.zero_()
        dst = PlanesVol.empty(B, 32, 2, 32, 64, 208, dev)
        for _ in range(2):
            ops.trilinear_ac(src, 0, 32, dst, 0)          # down-sample (cells 0, 11)
        del src, dst
        src = PlanesVol.empty(B, 32, 2, 32, 64, 208, dev); src.t.zero_()
        dst = PlanesVol.empty(B, 32, 2, 64, 128, 416, dev)
        for _ in range(2):
            ops.trilinear_ac(src, 0, 32, dst, 0)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
