"""BASELINE.json configs[4]: training forward+backward, SceneFlow crop 288x576, batch 4 per GPU, gradients averaged
with one NCCL all-reduce (torchrun for N > 1).  Prints one JSON line (ms per step, pairs/s)."""
import argparse, contextlib, io, json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from leastereo_b200 import LEAStereo, default_args
from leastereo_b200.sharding import allreduce_gradients
from leastereo_b200.pipeline import FlatAdam, masked_smooth_l1_loss

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4); ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=5); ap.add_argument("--h", type=int, default=288)
    ap.add_argument("--w", type=int, default=576); ap.add_argument("--conv", default="tc")
    ap.add_argument("--torch-optim", action="store_true", help="torch smooth_l1 + torch.optim.Adam instead of our kernels")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        model = LEAStereo(default_args(192), dev).to(dev).train()
    model.engine_options = {"planes": 2, "conv": a.conv}
    if a.torch_optim:
        opt = torch.optim.Adam(model.parameters(), lr=1e-3, betas=(0.9, 0.999))      # train.py:76
    else:
        opt = FlatAdam(model.parameters(), lr=1e-3, betas=(0.9, 0.999))              # one launch over the flat buffer
    g = torch.Generator().manual_seed(1 + rank)
    left = torch.randn(a.batch, 3, a.h, a.w, generator=g).to(dev)
    right = torch.randn(a.batch, 3, a.h, a.w, generator=g).to(dev)
    target = (torch.rand(a.batch, a.h, a.w, generator=g) * 96).to(dev)
    params = [p for p in model.parameters()]

    def step():
        if a.torch_optim:
            opt.zero_grad(set_to_none=True)
        else:
            opt.zero_grad()
        disp = model(left, right)
        if a.torch_optim:
            mask = (target < 192) & (target > 0.001)
            loss = torch.nn.functional.smooth_l1_loss(disp[mask], target[mask])
        else:
            loss = masked_smooth_l1_loss(disp, target, 192)                             # train.py:116-118,157
        loss.backward()
        if a.torch_optim:
            allreduce_gradients(params, world)
        else:
            opt.allreduce(world)                                                        # the flat bucket itself
        opt.step()
        return loss

    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    if dist: dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss = step()
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev, dtype=torch.float64)
    if dist: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"metric": "training fwd+bwd+Adam, SceneFlow crop %dx%d D=192, batch %d/GPU" % (a.h, a.w, a.batch),
                          "n_gpus": world, "ms_per_step": float(ms), "pairs_per_s": a.batch * world / (float(ms) / 1e3),
                          "loss": float(loss.detach()), "optimizer": "torch" if a.torch_optim else "FlatAdam + masked_smooth_l1 kernels", "conv": a.conv, "max_mem_GB": torch.cuda.max_memory_allocated() / 2**30}))
    if dist: dist.barrier(); dist.destroy_process_group()

if __name__ == "__main__":
    main()
