"""Host-side logic of the training-step engine on CPU: flat arenas keep values / physical order, and the
data-parallel gradient averaging is exercised with world_size 2 over gloo (no GPU)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from text_segmentation_image_inpainting_b200.engine import FlatParams, TrainStep


class Tiny(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.conv = torch.nn.Conv2d(3, 5, 3, bias=True)
        self.conv.weight.data = self.conv.weight.data.contiguous(memory_format=torch.channels_last)
        self.bn = torch.nn.BatchNorm2d(5)
        self.frozen = torch.nn.Conv2d(1, 1, 3, bias=False)
        self.frozen.weight.requires_grad = False


def test_flat_params_views_preserve_values_and_layout():
    torch.manual_seed(0)
    net = Tiny()
    before = {k: v.clone() for k, v in net.state_dict().items()}
    fp = FlatParams(net)
    after = net.state_dict()
    for k in before:
        assert torch.equal(before[k], after[k]), k
    w = net.conv.weight
    assert w.is_contiguous(memory_format=torch.channels_last) and w.grad.stride() == w.stride()
    assert w.data_ptr() == fp.flat_p.data_ptr() and w.grad.data_ptr() == fp.flat_g.data_ptr()
    assert all(o % 4 == 0 for o in fp.offsets) and fp.true_numel == 3 * 5 * 9 + 5 + 5 + 5
    assert not any(p is net.frozen.weight for p in fp.params)
    # autograd accumulates into the arena in place
    net(torch.randn(2, 3, 8, 8)) if False else None
    y = net.bn(torch.nn.functional.conv2d(torch.randn(2, 3, 8, 8), net.conv.weight, net.conv.bias))
    y.sum().backward()
    assert w.grad.data_ptr() == fp.flat_g.data_ptr() and float(fp.flat_g.abs().sum()) > 0


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(rank)                                        # replicas built from DIFFERENT seeds ...
    net = Tiny()
    net.bn.running_mean.fill_(float(rank + 1))
    ts = TrainStep(net, process_group=dist.group.WORLD, use_graph=False, bucket_mb=1)
    # ... must adopt rank 0's parameters and buffers at start-up (TrainStep._sync_replicas)
    gathered = [torch.empty_like(ts.flat.flat_p) for _ in range(world)]
    dist.all_gather(gathered, ts.flat.flat_p)
    same = all(torch.equal(gathered[0], g) for g in gathered) and float(net.bn.running_mean[0]) == 1.0
    ts.bucket_elems = 64                                           # force several buckets
    ts.flat.flat_g.copy_(torch.arange(ts.flat.numel, dtype=torch.float32) * (rank + 1))
    ts._allreduce()
    # ranks exchange the SUM; the 1/world of the mean is folded into the optimiser kernel (TrainStep.grad_scale)
    expect = torch.arange(ts.flat.numel, dtype=torch.float32) * (sum(range(1, world + 1)) / world)
    out[rank] = bool(torch.allclose(ts.flat.flat_g * ts.grad_scale, expect)) and ts.grad_scale == 1.0 / world and same
    dist.destroy_process_group()


def test_gradient_allreduce_averages_over_ranks_gloo_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}


def test_gradient_buckets_partition_the_arena_at_parameter_boundaries():
    torch.manual_seed(0)
    net = torch.nn.Sequential(*[torch.nn.Conv2d(8, 8, 3) for _ in range(6)])
    for m in net:
        m.weight.data = m.weight.data.contiguous(memory_format=torch.channels_last)
    ts = TrainStep(net, use_graph=False)
    ts.bucket_elems = 1000
    ts._make_buckets()
    assert ts.buckets[0][0] == 0 and ts.buckets[-1][1] == ts.flat.numel
    assert all(a[1] == b[0] for a, b in zip(ts.buckets, ts.buckets[1:]))
    assert sorted(i for _, _, mem in ts.buckets for i in mem) == list(range(len(ts.flat.params)))
    assert all(s in ts.flat.offsets for s, _, _ in ts.buckets) and len(ts.buckets) > 1
    # every parameter reports exactly once; a bucket is exchanged when its last member reports
    launched = []
    ts._comm_stream = None
    ts._launch_bucket = lambda b: launched.append(b)
    ts._hooks_installed = True
    ts._arm_overlap(True)
    for i in reversed(range(len(ts.flat.params))):
        ts._param_ready(i)
        ts._param_ready(i)
    assert launched == list(reversed(range(len(ts.buckets))))
