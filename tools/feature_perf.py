"""Timing of the stock-PyTorch 2D feature net (producer of the hot path's input) under different cuDNN settings."""
import contextlib, io, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from leastereo_b200 import LEAStereo, default_args

def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters

torch.manual_seed(0)
with contextlib.redirect_stdout(io.StringIO()):
    m = LEAStereo(default_args(192), "cuda").cuda().eval()
x = torch.randn(2, 3, 384, 1248, device="cuda")
with torch.no_grad():
    for bench in (False, True):
        for cl in (False, True):
            xx = x.to(memory_format=torch.channels_last) if cl else x
            f = m.feature.to(memory_format=torch.channels_last) if cl else m.feature.to(memory_format=torch.contiguous_format)
            with torch.backends.cudnn.flags(enabled=True, benchmark=bench, allow_tf32=False):
                ms = timeit(lambda: f(xx))
                g = torch.cuda.CUDAGraph()
                s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s):
                    f(xx)
                torch.cuda.current_stream().wait_stream(s)
                with torch.cuda.graph(g):
                    y = f(xx)
                msg = timeit(lambda: g.replay())
            print("benchmark=%s channels_last=%s eager %.3f ms  graph %.3f ms" % (bench, cl, ms, msg))
