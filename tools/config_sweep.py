"""Throughput of the other BASELINE.json inference configs (device-resident, CUDA events, through LEAStereo.forward = the module-owned CUDA graph):
   configs[1] SceneFlow 576x960 D=192 batch 8, configs[2] KITTI 384x1248 batch sweep, configs[3] Middlebury half-res
   1008x1512 D=408.  Prints one JSON line per case; bench.py remains the contract benchmark (KITTI, 4 pairs/step)."""
import contextlib, io, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from leastereo_b200 import LEAStereo, default_args  # noqa: E402


def run(name, H, W, maxdisp, B, steps=5, warmup=4):     # the module captures its CUDA graph on the third call
    dev = torch.device("cuda:0")
    torch.cuda.empty_cache(); torch.cuda.reset_peak_memory_stats()
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        model = LEAStereo(default_args(maxdisp=maxdisp, cuda=True), dev).to(dev).eval()
    model.engine_options = {"assume_frozen": True}
    g = torch.Generator().manual_seed(1)
    left = torch.randn(B, 3, H, W, generator=g).to(dev); right = torch.randn(B, 3, H, W, generator=g).to(dev)
    with torch.no_grad():
        for _ in range(warmup):
            d = model(left, right)
        torch.cuda.synchronize()
        torch.cuda._sleep(int(2e7))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            d = model(left, right)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    out = {"config": name, "H": H, "W": W, "maxdisp": maxdisp, "batch": B, "ms_per_step": round(ms, 3),
           "pairs_per_s": round(B / ms * 1e3, 2), "finite": bool(torch.isfinite(d).all()),
           "max_mem_GB": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1)}
    print(json.dumps(out), flush=True)
    del model, left, right, d
    torch.cuda.empty_cache(); torch.cuda.reset_peak_memory_stats()
    return out


if __name__ == "__main__":
    res = [run("SceneFlow crop (configs[0] shape)", 288, 576, 192, 1),
           run("SceneFlow full frame (configs[1])", 576, 960, 192, 8),
           run("KITTI batch 1 (configs[2])", 384, 1248, 192, 1),
           run("KITTI batch 4 (configs[2])", 384, 1248, 192, 4),
           run("KITTI batch 16 (configs[2])", 384, 1248, 192, 16),
           run("Middlebury half-res (configs[3])", 1008, 1512, 408, 1),
           run("Middlebury half-res, batch 2", 1008, 1512, 408, 2)]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "config_sweep.json"), "w"), indent=1)
