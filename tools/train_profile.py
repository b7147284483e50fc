"""Development aid: top CUDA kernels of one training step (torch.profiler)."""
import contextlib, io, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from leastereo_b200 import LEAStereo, default_args
from torch.profiler import profile, ProfilerActivity

dev = torch.device("cuda:0")
torch.manual_seed(0)
with contextlib.redirect_stdout(io.StringIO()):
    model = LEAStereo(default_args(192), dev).to(dev).train()
model.engine_options = {"planes": 2, "conv": "tc"}
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

for _ in range(5):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=70))
