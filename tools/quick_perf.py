"""Quick device timings of the individual kernels at full KITTI size (CUDA events, after warm-up).  Development aid;
bench.py is the contract benchmark."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from leastereo_b200.kernels import get_ops, PlanesVol, lea_tc_opts  # noqa: E402


def timeit(fn, iters=5, warm=2):
    """Median device time of fn's launches.  A long spin kernel is queued first so that the host runs ahead of the
    GPU and the event pairs bracket back-to-back GPU work (otherwise small kernels measure the host launch path)."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    torch.cuda._sleep(int(4e6 * (iters + 1)))
    evs = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2]


def main():
    ops = get_ops()
    dev = torch.device("cuda:0")
    out = {}
    B, C, H3, W3, D3, maxdisp = 1, 32, 128, 416, 64, 192
    x = torch.randn(B, C, H3, W3, device=dev); y = torch.randn(B, C, H3, W3, device=dev)
    cost = torch.empty((B, 2 * C, D3, H3, W3), device=dev)
    import ctypes as Ct
    def cv():
        ops._check(ops.lib.lea_cost_volume_f32(x.data_ptr(), y.data_ptr(), cost.data_ptr(), B, C, H3, W3, D3,
                                               ops._stream(x)))
    ms = timeit(cv)
    nbytes = 4 * (2 * C * H3 * W3 + 2 * C * D3 * H3 * W3) * B
    out["cost_volume_f32"] = {"ms": ms, "GBps": nbytes / ms / 1e6}
    vol = PlanesVol.empty(B, 2 * C, 2, D3, H3, W3, dev)
    ms = timeit(lambda: ops.cost_volume_planes(x, y, maxdisp, 2, out=vol))
    out["cost_volume_planes"] = {"ms": ms, "GBps": nbytes / ms / 1e6}
    mat = torch.randn(B, D3, H3, W3, device=dev) * 3
    ms = timeit(lambda: ops.disp_head(mat, maxdisp))
    out["disp_head"] = {"ms": ms, "GBps": 4 * (D3 * H3 * W3 + 9 * H3 * W3) * B / ms / 1e6,
                        "Gexp_per_s": maxdisp * 9 * H3 * W3 * B / ms / 1e6}
    # convs at KITTI shapes: (name, c_in, c_out, k, spatial)
    shapes = [("stem0", 64, 32, 3, (64, 128, 416)), ("stem1", 32, 32, 3, (64, 128, 416)),
              ("conv1", 128, 64, 3, (32, 64, 208)), ("L1op", 16, 16, 3, (32, 64, 208)),
              ("L2op", 32, 32, 3, (16, 32, 104)), ("L0op", 8, 8, 3, (64, 128, 416)),
              ("last_3", 32, 1, 3, (64, 128, 416)), ("pp128_16", 128, 16, 1, (32, 64, 208)),
              ("pp64_8", 64, 8, 1, (64, 128, 416))]
    modes = sys.argv[1:] or ["simt", "tc"]
    for name, ci, co, k, sp in shapes:
        src = PlanesVol.empty(B, ci, 2, *sp, dev)
        src.t.copy_(torch.randn(src.t.shape, device=dev).bfloat16() * 0.1)
        w = torch.randn(co, ci, k, k, k, device=dev) * 0.05
        scale = torch.ones(co, device=dev); shift = torch.zeros(co, device=dev)
        flops = 2.0 * B * sp[0] * sp[1] * sp[2] * ci * co * k ** 3
        if co % 8 == 0:
            dst = PlanesVol.empty(B, co, 2, *sp, dev)
            p = ops.make_conv(src, 0, ci, co, k, scale, shift, True, dst=dst)
        else:
            o = torch.empty((B, co) + sp, device=dev)
            p = ops.make_conv(src, 0, ci, co, k, None, None, False, dst_f32=o)
        for mode in modes:
            try:
                if mode == "simt":
                    ms = timeit(lambda: ops.conv3d_simt(p, w, w), iters=3, warm=1)
                else:
                    img = ops.pack_weights_tc(w, 2)
                    opts = lea_tc_opts()
                    opts.accum_split = int(os.environ.get("LEA_SPLIT", "0"))
                    opts.acc_sets = int(os.environ.get("LEA_SETS", "0"))
                    ms = timeit(lambda: ops.conv3d_tc(p, img, opts, w), iters=3, warm=1)
                out["conv_%s_%s" % (mode, name)] = {"ms": ms, "TFLOPs": flops / ms / 1e9}
            except Exception as e:  # noqa: BLE001
                out["conv_%s_%s" % (mode, name)] = {"error": str(e)[:200]}
                break
        del src
    for k_, v in out.items():
        print(k_, json.dumps(v))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "quick_perf_%s.json" % "_".join(modes)), "w"), indent=1)


if __name__ == "__main__":
    main()
