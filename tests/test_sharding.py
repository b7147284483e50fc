"""CPU tests (-m "not gpu") of the multi-GPU partitioning logic with the gloo backend, world_size 2 and 3.
The forward function is the oracle here (the CUDA path needs a GPU); what is under test is the sharding/gather code
the bench and a multi-GPU caller use: pair order, ragged and empty shards."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import load_golden, golden_state_dict
from leastereo_b200.sharding import shard_range, sharded_inference


def test_shard_range_properties():
    for n in range(0, 20):
        for ws in range(1, 9):
            blocks = [shard_range(n, r, ws) for r in range(ws)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_pairs, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from oracle import leastereo_oracle as O
    g = load_golden("cal_b2_24x48_d24")
    sd = golden_state_dict(g)
    left = torch.from_numpy(g["left"]).repeat((n_pairs + 1) // 2, 1, 1, 1)[:n_pairs]
    right = torch.from_numpy(g["right"]).repeat((n_pairs + 1) // 2, 1, 1, 1)[:n_pairs]
    left = left + 0.01 * torch.arange(n_pairs, dtype=torch.float32).view(-1, 1, 1, 1)   # make every pair distinct

    def fwd(l, r):
        return O.leastereo_forward(sd, l, r, int(g["maxdisp"]))

    out = sharded_inference(fwd, left, right, rank, world)
    if rank == 0:
        torch.save({"sharded": out, "full": fwd(left, right)}, out_path)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_pairs", [(2, 3), (3, 2)])
def test_sharded_inference_matches_single_process(tmp_path, world, n_pairs):
    out_path = str(tmp_path / "out.pt")
    mp.spawn(_worker, args=(world, _free_port(), n_pairs, out_path), nprocs=world, join=True)
    res = torch.load(out_path)
    assert res["sharded"].shape == res["full"].shape
    # BN is in eval mode, so a pair's result does not depend on its batch-mates: sharding must be exact up to the
    # fp32 re-ordering noise of the CPU convs, which pick different algorithms for different batch sizes
    # (that noise floor is max 9e-3 px, mean 4e-4 px on this network)
    diff = (res["sharded"] - res["full"]).abs()
    assert float(diff.max()) <= 0.05 and float(diff.mean()) <= 1e-3


def _grad_worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from leastereo_b200.sharding import allreduce_gradients
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(3, 4)), torch.nn.Parameter(torch.zeros(5)),
              torch.nn.Parameter(torch.zeros(2, 2)), torch.nn.Parameter(torch.zeros(7), requires_grad=False)]
    params[0].grad = torch.full((3, 4), float(rank + 1))
    params[1].grad = torch.arange(5.0) * (rank + 1)
    # params[2] never receives a gradient (like the reference's unused heads); params[3] is frozen
    n = allreduce_gradients(params, world)
    if rank == 0:
        torch.save({"n": n, "g0": params[0].grad, "g1": params[1].grad, "g2": params[2].grad}, out_path)
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_allreduce_flat_bucket(tmp_path):
    out_path = str(tmp_path / "g.pt")
    mp.spawn(_grad_worker, args=(2, _free_port(), out_path), nprocs=2, join=True)
    res = torch.load(out_path)
    assert res["n"] == 12 + 5 + 4
    assert torch.allclose(res["g0"], torch.full((3, 4), 1.5))
    assert torch.allclose(res["g1"], torch.arange(5.0) * 1.5)
    assert res["g2"] is None


def _flat_adam_worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import shutil
    from test_emu_kernels import EMU_LIB
    from leastereo_b200.kernels import Ops
    from leastereo_b200.pipeline import FlatAdam
    ops = Ops(EMU_LIB, require_device_build=False)           # CPU emulation of the Adam kernel (test infrastructure)
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.ones(3, 4)), torch.nn.Parameter(torch.ones(5))]
    opt = FlatAdam(params, lr=0.1, ops=ops)
    opt.zero_grad()
    params[0].grad.add_(float(rank + 1))                      # accumulates in the flat bucket, like autograd
    params[1].grad = torch.arange(5.0) * (rank + 1)           # a foreign .grad tensor
    opt.allreduce(world)
    g = opt.grad.clone()
    opt.step()
    if rank == 0:
        torch.save({"g": g, "p0": params[0].detach().clone(), "p1": params[1].detach().clone()}, out_path)
    dist.barrier()
    dist.destroy_process_group()


def test_flat_adam_allreduce(tmp_path):
    """FlatAdam.allreduce: the flat gradient bucket is the all-reduce buffer; both ranks then take the same step."""
    import shutil
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    from test_emu_kernels import EMU_LIB
    if not os.path.exists(EMU_LIB):
        import __graft_entry__ as ge
        ge.build()
    out_path = str(tmp_path / "fa.pt")
    mp.spawn(_flat_adam_worker, args=(2, _free_port(), out_path), nprocs=2, join=True)
    res = torch.load(out_path)
    assert torch.allclose(res["g"][:12], torch.full((12,), 1.5))
    assert torch.allclose(res["g"][12:], torch.arange(5.0) * 1.5)
    # first Adam step moves every parameter with a non-zero gradient by lr against the gradient's sign
    assert torch.allclose(res["p0"], torch.full((3, 4), 0.9), atol=1e-5)
    assert torch.allclose(res["p1"][1:], torch.full((4,), 0.9), atol=1e-5) and float(res["p1"][0]) == 1.0
