"""Soak: many replays of the KITTI step (CUDA graph) + a sweep of odd shapes through the whole module; checks that
results stay bit-identical across replays and finite across shapes (catches barrier-phase / race bugs)."""
import contextlib, io, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from leastereo_b200 import LEAStereo, default_args

dev = torch.device("cuda:0")
torch.manual_seed(0)
with contextlib.redirect_stdout(io.StringIO()):
    model = LEAStereo(default_args(maxdisp=192, cuda=True), dev).to(dev).eval()
model.engine_options = {"assume_frozen": True}
g = torch.Generator().manual_seed(1)
left = torch.randn(4, 3, 384, 1248, generator=g).to(dev); right = torch.randn(4, 3, 384, 1248, generator=g).to(dev)
with torch.no_grad():
    ref = model(left, right).clone()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        model(left, right)
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = model(left, right)
    t0 = time.time()
    n = int(os.environ.get("SOAK_STEPS", "300"))
    bad = 0
    for i in range(n):
        graph.replay()
        if i % 25 == 24:
            torch.cuda.synchronize()
            if not torch.equal(out, ref):
                bad += 1
    torch.cuda.synchronize()
    print("replays", n, "mismatching checks", bad, "time %.1f s" % (time.time() - t0), "finite", bool(torch.isfinite(out).all()))
    # odd shapes through the eager path (new plans each time)
    for (B, H, W, md) in [(1, 96, 192, 48), (2, 120, 264, 96), (1, 192, 384, 192), (3, 72, 216, 60), (1, 240, 600, 192),
                          (1, 384, 1248, 96), (2, 288, 576, 192), (1, 168, 456, 129)]:
        with contextlib.redirect_stdout(io.StringIO()):
            m2 = LEAStereo(default_args(maxdisp=md, cuda=True), dev).to(dev).eval()
        l = torch.randn(B, 3, H, W, generator=g).to(dev); r = torch.randn(B, 3, H, W, generator=g).to(dev)
        try:
            d1 = m2(l, r); d2 = m2(l, r)
            torch.cuda.synchronize()
            print((B, H, W, md), tuple(d1.shape), "finite", bool(torch.isfinite(d1).all()), "repeatable", bool(torch.equal(d1, d2)))
        except Exception as e:  # noqa: BLE001
            print((B, H, W, md), "raised", type(e).__name__, str(e)[:120])
        del m2
