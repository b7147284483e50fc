"""Soak of the module-owned graphs: many eval calls through LEAStereo.forward (graph replays) with changing inputs checked
against eager results, odd shapes (flat mode, ragged tiles) called often enough to capture, then a training loop on the
captured forward / backward graphs; watches device memory for growth."""
import contextlib, io, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from leastereo_b200 import LEAStereo, default_args
from leastereo_b200.pipeline import FlatAdam, masked_smooth_l1_loss

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(1)

def make(md):
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        return LEAStereo(default_args(maxdisp=md, cuda=True), dev).to(dev)

with torch.no_grad():
    for (B, H, W, md) in [(2, 384, 1248, 192), (1, 96, 192, 48), (3, 72, 216, 72), (1, 168, 456, 144), (1, 240, 600, 192)]:   # 1/3-resolution dims multiples of 8, as the reference needs
        m = make(md).eval()
        ins = [(torch.randn(B, 3, H, W, generator=g).to(dev), torch.randn(B, 3, H, W, generator=g).to(dev)) for _ in range(3)]
        m.engine_options = {"cuda_graph": False}
        want = [m(l, r).clone() for l, r in ins]
        m.engine_options = {}
        bad = 0
        mem0 = None
        for i in range(60):
            l, r = ins[i % 3]
            out = m(l, r)
            if not torch.equal(out, want[i % 3]):
                bad += 1
            if i == 10:
                torch.cuda.synchronize(); mem0 = torch.cuda.memory_allocated()
        torch.cuda.synchronize()
        print((B, H, W, md), "mismatches", bad, "finite", bool(torch.isfinite(out).all()),
              "memory growth MB", round((torch.cuda.memory_allocated() - mem0) / 2 ** 20, 2), flush=True)
        del m

m = make(192).train()
opt = FlatAdam(m.parameters(), lr=1e-3)
l = torch.randn(2, 3, 192, 384, generator=g).to(dev); r = torch.randn(2, 3, 192, 384, generator=g).to(dev)
t = (torch.rand(2, 192, 384, generator=g) * 90).to(dev)
losses, mem0 = [], None
for i in range(40):
    opt.zero_grad()
    loss = masked_smooth_l1_loss(m(l, r), t, 192)
    loss.backward()
    opt.step()
    losses.append(float(loss))
    if i == 8:
        torch.cuda.synchronize(); mem0 = torch.cuda.memory_allocated()
torch.cuda.synchronize()
print("train: loss", round(losses[0], 4), "->", round(losses[-1], 4), "finite", all(x == x for x in losses),
      "decreasing", losses[-1] < losses[0], "memory growth MB", round((torch.cuda.memory_allocated() - mem0) / 2 ** 20, 2))
