"""CPU tests (world_size 2, gloo) of the N > 1 host logic in human_body_reconstruction_b200/dist.py:
ray sharding, the flat-gradient all-reduce launched from inside backward (DDP-style end-of-backward wait),
slab sharding of the density grid and the all-gather of marching-cubes counts.  The kernels themselves need a GPU;
here stand-in modules speak the same dist._GradExchange protocol HashEncoder / MLP_3D use, including several passes
through one module inside one backward (hierarchical render) and gradient accumulation across backward calls."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as tdist
import torch.multiprocessing as mp

from human_body_reconstruction_b200 import dist as hdist
from oracle import port


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _FlatGradModule(torch.nn.Module):
    """Stand-in for HashEncoder / MLP_3D: per-"level" parameters that are views of one flat buffer, speaking the modules'
    side of the dist._GradExchange protocol (note_forward / enter_backward / publish, gradients never returned to
    autograd while a reducer is attached)."""

    def __init__(self, levels, width, chunks=1):
        super().__init__()
        g = torch.Generator().manual_seed(7)
        self.flat = torch.randn(levels, width, generator=g)
        self.levels = torch.nn.ParameterList([torch.nn.Parameter(self.flat[i]) for i in range(levels)])
        self._dp = None
        self._grad_chunks = chunks

    def _dp_template(self):
        return self.flat

    def _dp_param_views(self, buf):
        return [(p, buf[i]) for i, p in enumerate(self.levels)]

    def forward(self, x):
        mod = self

        class Fn(torch.autograd.Function):
            @staticmethod
            def forward(ctx, x, *w):
                ctx.save_for_backward(x)
                if mod._dp is not None:
                    mod._dp.note_forward(mod)
                return x @ mod.flat.detach().T                       # (N, levels)

            @staticmethod
            def backward(ctx, dy):
                (x,) = ctx.saved_tensors
                L = mod.flat.shape[0]
                if mod._dp is None:
                    gflat = dy.T @ x
                    return (None,) + tuple(gflat[i] for i in range(L))
                gflat, last = mod._dp.enter_backward(mod)
                nch = mod._dp.chunks(mod) if last else 1
                step = -(-L // nch)
                for l0 in range(0, L, step):                         # level chunks, published as they complete
                    l1 = min(L, l0 + step)
                    gflat[l0:l1] += dy[:, l0:l1].T @ x
                    if last:
                        mod._dp.publish(mod, gflat[l0:l1])
                return (None,) * (L + 1)

        return Fn.apply(x, *self.levels)


def _worker(rank, world, port_no, tmp):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no))
    r, w = hdist.init_from_env("gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(0)                                             # identical parameters on every rank
    n_rays, width, levels = 64, 5, 6
    g = torch.Generator().manual_seed(3)
    x_all = torch.randn(n_rays, width, generator=g)
    x2_all = torch.randn(n_rays, width, generator=g)
    gt_all = torch.randn(n_rays, levels, generator=g)

    def loss_of(mod, sl, passes):
        # passes == 2: the module is evaluated twice inside one loss, like the coarse + fine render of
        # vol_renderer.py:220,242 -- two autograd nodes feeding the same flat gradient
        loss = torch.nn.functional.mse_loss(mod(x_all[sl]), gt_all[sl])   # mean over the LOCAL batch
        if passes == 2:
            loss = loss + torch.nn.functional.mse_loss(mod(x2_all[sl]), gt_all[sl])
        return loss

    for chunks, passes in ((1, 1), (3, 1), (1, 2), (3, 2)):
        mod = _FlatGradModule(levels, width, chunks=chunks)
        if chunks == 1:
            reducer = hdist.GradAllReduce(mod)
        else:       # the front door picks the NCCL/gloo call path when the ranks are not CUDA ranks (no peer memory here)
            reducer = hdist.attach_grad_allreduce(mod, _FlatGradModule(levels, width))
            assert isinstance(reducer, hdist.GradAllReduce)
        sl = hdist.shard_rays(n_rays, rank, world)
        loss_of(mod, sl, passes).backward()
        got = torch.stack([p.grad for p in mod.levels])
        # single-process gradient of the concatenated batch
        ref = _FlatGradModule(levels, width)
        loss_of(ref, slice(None), passes).backward()
        want = torch.stack([p.grad for p in ref.levels])
        assert torch.allclose(got, want, rtol=1e-5, atol=1e-6), (rank, chunks, passes, (got - want).abs().max())
        assert reducer.bytes_reduced == levels * width * 4, "every buffer is reduced exactly once per backward"
        # a second backward WITHOUT clearing .grad accumulates (the persistent buffer is not handed out twice)
        loss_of(mod, sl, passes).backward()
        got2 = torch.stack([p.grad for p in mod.levels])
        assert torch.allclose(got2, 2 * want, rtol=1e-5, atol=1e-6), (rank, chunks, passes)
        assert reducer.sessions == 2
        reducer.remove()
        assert mod._dp is None

    # density-grid slabs: disjoint, ordered, covering; per-slab crossing-edge counts add up to the whole grid's
    res = 11
    i0, i1 = hdist.slab_range(res, rank, world)
    rng = np.random.default_rng(5)
    dens = rng.normal(30.0, 5.0, size=(res, res, res)).astype(np.float32)
    whole = port.mc_crossing_edges(dens, 30.0)
    mine = port.mc_crossing_edges(dens, 30.0, i_begin=i0, i_end=i1)
    counts = hdist.allgather_counts([mine, i1 - i0])
    assert len(counts) == world and counts[rank] == [mine, i1 - i0]
    assert sum(c[0] for c in counts) == whole
    assert sum(c[1] for c in counts) == res
    if rank == 0:
        with open(os.path.join(tmp, "ok"), "w") as f:
            f.write("ok")
    tdist.barrier()
    tdist.destroy_process_group()


def test_two_rank_gradient_allreduce_and_slabs(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert (tmp_path / "ok").read_text() == "ok"


@pytest.mark.parametrize("n,world", [(4096, 8), (10, 3), (7, 8), (1 << 20, 8)])
def test_shard_rays_partition(n, world):
    seen = []
    for r in range(world):
        s = hdist.shard_rays(n, r, world)
        seen += list(range(n))[s]
    assert seen == list(range(n))


@pytest.mark.parametrize("res,world", [(512, 8), (256, 3), (5, 8), (11, 2)])
def test_slab_range_partition(res, world):
    edges = [hdist.slab_range(res, r, world) for r in range(world)]
    assert edges[0][0] == 0 and edges[-1][1] == res
    for (a0, a1), (b0, b1) in zip(edges, edges[1:]):
        assert a1 == b0 and a0 <= a1
    sizes = [b - a for a, b in edges]
    assert max(sizes) - min(sizes) <= 1


def test_level_chunks_boundaries():
    """dist._GradExchange.level_chunks: a count gives equal level ranges, a list gives inner boundaries; both cover [0, L)."""
    from human_body_reconstruction_b200.dist import _GradExchange

    class Fake(_GradExchange):
        def __init__(self, c):
            self._c = c

        def chunks(self, m):
            return self._c

    assert Fake(1).level_chunks(None, 16) == [(0, 16)]
    assert Fake(2).level_chunks(None, 16) == [(0, 8), (8, 16)]
    assert Fake(3).level_chunks(None, 16) == [(0, 6), (6, 12), (12, 16)]
    assert Fake([4, 8, 12, 14]).level_chunks(None, 16) == [(0, 4), (4, 8), (8, 12), (12, 14), (14, 16)]
    assert Fake([12, 4, 0, 16, 99]).level_chunks(None, 16) == [(0, 4), (4, 12), (12, 16)]
    for c in (1, 2, 5, 16, 40, [3], [1, 2, 3]):
        ch = Fake(c).level_chunks(None, 16)
        assert ch[0][0] == 0 and ch[-1][1] == 16 and all(a[1] == b[0] for a, b in zip(ch, ch[1:]))


def test_stream_plan_and_level_bounds_host_logic():
    """Host side of the streamed exchange: the base exchange never streams; PeerGradAllReduce.stream_plan streams only the
    encoder's last table backward, only when `streamed`, only below CHUNK_MAX_POINTS and only with more than one chunk;
    ops._bounds turns a chunk list into the C ABI's level_bounds and rejects lists that do not partition the levels."""
    from human_body_reconstruction_b200 import ops
    from human_body_reconstruction_b200.dist import PeerGradAllReduce, _GradExchange

    class Base(_GradExchange):
        def __init__(self):
            pass

    assert Base().stream_plan(object(), 16, 1000) is None

    enc, other = object(), object()

    class Fake(PeerGradAllReduce):
        def __init__(self, streamed, chunks):
            self.streamed, self.overlap, self._nchunks, self._enc = streamed, True, chunks, enc

    assert Fake(True, 4).stream_plan(enc, 16, 524288) == [(0, 4), (4, 8), (8, 12), (12, 16)]
    assert Fake(True, [6, 10, 13]).stream_plan(enc, 16, 524288) == [(0, 6), (6, 10), (10, 13), (13, 16)]
    assert Fake(False, 4).stream_plan(enc, 16, 524288) is None                        # launch pair per chunk
    assert Fake(True, 4).stream_plan(other, 16, 524288) is None                       # not the encoder
    assert Fake(True, 4).stream_plan(enc, 16, Fake.CHUNK_MAX_POINTS + 1) is None      # long steps are not chunked
    assert Fake(True, 1).stream_plan(enc, 16, 524288) is None                         # a single chunk: nothing to stream
    assert Fake(True, 4).stream_plan(enc, 16, 0) is None

    b, n = ops._bounds([(0, 4), (4, 8), (8, 16)], 16)
    assert n == 3 and list(b) == [0, 4, 8, 16]
    for bad in ([(0, 4), (5, 16)], [(0, 4), (4, 12)], [(1, 16)], [(0, 8), (8, 8), (8, 16)], [(0, 8), (4, 16)]):
        with pytest.raises(ValueError):
            ops._bounds(bad, 16)


def _small_param_worker(rank, world, port_no, tmp):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no))
    tdist.init_process_group("gloo", rank=rank, world_size=world)
    from human_body_reconstruction_b200.helper import VarModel
    var = VarModel()
    with torch.no_grad():
        var.b.fill_(0.5 + rank)                                        # every process initialised its own
    assert hdist.attach_small_params(var.parameters()) == [var.b]
    assert hdist.attach_small_params(var.parameters()) == []           # idempotent: one hook per parameter
    b0 = float(var.b)
    x = torch.linspace(-1, 1, 8) * (rank + 1)
    for _ in range(2):                                                 # two backward passes: the hook fires on each
        var.b.grad = None
        var(x).sum().backward()
    torch.save((b0, var.b.grad.clone()), os.path.join(tmp, f"b{rank}.pt"))
    tdist.barrier()
    tdist.destroy_process_group()


def test_small_params_broadcast_and_gradient_average(tmp_path):
    """dist.attach_small_params (SDF mode's VarModel.b under launch_rank.py): rank 0's value everywhere, gradient = mean
    over the ranks' local gradients, identical on both ranks."""
    mp.spawn(_small_param_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    (b0, g0), (b1, g1) = (torch.load(os.path.join(tmp_path, f"b{r}.pt")) for r in range(2))
    assert b0 == 0.5 and b1 == 0.5
    b = torch.tensor(0.5, requires_grad=True)
    want = 0
    for r in range(2):
        x = torch.linspace(-1, 1, 8) * (r + 1)
        (g,) = torch.autograd.grad((1 / (1 + torch.exp(-x * b))).sum(), b)
        want = want + g / 2
    assert torch.equal(g0, g1) and torch.allclose(g0, want, rtol=1e-6)


def _auto_attach_worker(rank, world, port_no, tmp):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no))
    tdist.init_process_group("gloo", rank=rank, world_size=world)
    import human_body_reconstruction_b200 as h
    torch.manual_seed(100 + rank)                                      # the unchanged trainer does not seed: ranks differ
    enc = h.HashEncoder(N_min=16, N_max=64.0, L=4, F=2, T=64, dim=3, mu=torch.zeros(3), sigma=torch.tensor(1.0), device="cpu")
    mlp = h.MLP_3D(num_sig=2, num_col=2, L=4, F=2, d_view=24)
    var = h.helper.VarModel()
    with torch.no_grad():
        var.b.fill_(0.25 + rank)
    red = hdist.auto_attach(enc, mlp, extra=var.parameters())
    again = hdist.auto_attach(enc, mlp, extra=var.parameters())        # second call: nothing new, same exchange
    torch.save({"table": enc._flat_table().clone(), "mlp": mlp._flat_params().clone(), "b": float(var.b),
                "kind": type(red).__name__, "same": again is red, "attached": enc._dp is red and mlp._dp is red,
                "hook": getattr(var.b, "_hbr_dp_hook", None) is not None}, os.path.join(tmp, f"aa{rank}.pt"))
    tdist.barrier()
    tdist.destroy_process_group()


def test_auto_attach_broadcasts_rank0_parameters_and_takes_the_sdf_parameter_along(tmp_path):
    """dist.auto_attach (launch_rank.py's zero-edit route): every rank ends with rank 0's tables, MLP parameters and --
    SDF mode -- VarModel.b; the exchange is attached once (gloo ranks: the torch.distributed all-reduce flavour)."""
    mp.spawn(_auto_attach_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    a, b = (torch.load(os.path.join(tmp_path, f"aa{r}.pt")) for r in range(2))
    assert torch.equal(a["table"], b["table"]) and torch.equal(a["mlp"], b["mlp"]) and a["b"] == b["b"] == 0.25
    for r in (a, b):
        assert r["kind"] == "GradAllReduce" and r["same"] and r["attached"] and r["hook"]
