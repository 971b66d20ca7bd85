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
    torch.manual_seed(0)
    net = Tiny()
    ts = TrainStep(net, process_group=dist.group.WORLD, use_graph=False, bucket_mb=1)
    ts.bucket_elems = 64                                           # force several buckets
    ts.flat.flat_g.copy_(torch.arange(ts.flat.numel, dtype=torch.float32) * (rank + 1))
    ts._allreduce()
    expect = torch.arange(ts.flat.numel, dtype=torch.float32) * (sum(range(1, world + 1)) / world)
    out[rank] = bool(torch.allclose(ts.flat.flat_g, expect))
    dist.destroy_process_group()


def test_gradient_allreduce_averages_over_ranks_gloo_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}
