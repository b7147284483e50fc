"""Development aid: CUDA-event time of every node of the training plan (forward and backward), eager launches."""
import contextlib, io, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from leastereo_b200 import LEAStereo, default_args
from leastereo_b200 import training

dev = torch.device("cuda:0")
torch.manual_seed(0)
with contextlib.redirect_stdout(io.StringIO()):
    model = LEAStereo(default_args(192), dev).to(dev).train()
model.engine_options = {"planes": 2, "conv": "tc", "train_graph": False}
g = torch.Generator().manual_seed(1)
B, H, W = 4, 288, 576
left = torch.randn(B, 3, H, W, generator=g).to(dev); right = torch.randn(B, 3, H, W, generator=g).to(dev)
target = (torch.rand(B, H, W, generator=g) * 96).to(dev)

def step():
    model.zero_grad(set_to_none=True)
    disp = model(left, right)
    mask = (target < 192) & (target > 0.001)
    loss = torch.nn.functional.smooth_l1_loss(disp[mask], target[mask])
    loss.backward()

step(); step(); torch.cuda.synchronize()
plan = next(iter(training._TRAIN_PLANS.values()))
rec = []
for n in plan.nodes:
    for which in ("forward", "backward"):
        fn = getattr(n, which)
        def wrapped(fn=fn, n=n, which=which):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            rec.append((n.name, type(n).__name__, which, a, b))
        setattr(n, which, wrapped)
step(); torch.cuda.synchronize()
rows = [(nm, ty, wh, a.elapsed_time(b)) for nm, ty, wh, a, b in rec]
tot = {}
for nm, ty, wh, ms in rows:
    tot[(ty, wh)] = tot.get((ty, wh), 0.0) + ms
print(json.dumps({"by_type_ms": {"%s.%s" % k: round(v, 3) for k, v in sorted(tot.items())},
                  "total_ms": round(sum(r[3] for r in rows), 3)}))
for nm, ty, wh, ms in sorted(rows, key=lambda r: -r[3])[:60]:
    print("%-46s %-14s %-9s %8.1f us" % (nm, ty, wh, ms * 1e3))
