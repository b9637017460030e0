"""hbr_allreduce_peer through the C ABI on one device (world = 1: the kernel degenerates to an in-place scale; the
multi-rank exchange is covered by tests/test_gpu_multi.py on >= 2 GPUs)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_peer_region_single_rank_scale_and_ranges():
    from human_body_reconstruction_b200.peer import PeerRegion
    reg = PeerRegion(100003)                       # not a multiple of 4: the region pads itself
    assert reg.n % 4 == 0 and reg.n >= 100003 and reg.world == 1
    assert float(reg.tensor.abs().max()) == 0.0    # zero-filled at allocation
    src = torch.randn(reg.n, device="cuda")
    reg.tensor.copy_(src)
    reg.all_reduce(scale=0.5)
    assert torch.equal(reg.tensor, src * 0.5)
    reg.all_reduce(n=4096, scale=2.0, offset=1024, ctas=8)     # a sub-range (level chunk) leaves the rest alone
    want = src * 0.5
    want[1024:1024 + 4096] *= 2.0
    assert torch.equal(reg.tensor, want)
    assert not reg.timed_out()
    with pytest.raises(ValueError):
        reg.all_reduce(n=8, offset=2)
    reg.close()


def test_peer_abi_argument_errors():
    import ctypes as C
    from human_body_reconstruction_b200 import _lib
    lib = _lib.lib()
    one = (C.c_void_p * 1)(C.c_void_p(256))
    assert lib.hbr_allreduce_peer(one, one, None, 0, 9, 4, 1.0, 0, None, None) != 0        # world > HBR_MAX_PEERS
    assert lib.hbr_allreduce_peer(one, one, None, 0, 1, 6, 1.0, 0, None, None) != 0        # n % 4 != 0
    assert b"multiple of 4" in lib.hbr_last_error()


@pytest.mark.parametrize("form", ["pts", "rays", "rays_lm"])
def test_streamed_scatter_and_exchange_single_rank(form):
    """The streamed gradient exchange on one device (world = 1: the exchange degenerates to an in-place scale): ONE
    scatter-add launch that finishes the level chunks in order and counts its CTAs per chunk (hbr_hash_encode_bwd*_stream),
    ONE exchange launch on a side stream beside it that waits for each chunk's count (hbr_allreduce_peer_stream).  The
    result equals scale * the one-launch-per-range scatter-add (same reductions per entry; atomics: norm-wise), the
    counters end at the tile count, and a piece with need = 0 travels without a counter."""
    from human_body_reconstruction_b200 import ops
    from human_body_reconstruction_b200.peer import PeerRegion
    torch.manual_seed(3)
    L, F, T, R, S = 16, 2, 2 ** 14, 96, 40                                 # 3840 points: 30 tiles
    geom = ops.make_geom([-1.0, -1.1, -0.9], 3.7, [16.0 * 1.38 ** i for i in range(L)], L, F, T, 0)
    ro = torch.rand(R, 3, device="cuda") - 0.5
    rd = torch.nn.functional.normalize(torch.randn(R, 3, device="cuda"), dim=-1)
    t = torch.linspace(0.1, 1.5, S, device="cuda")
    dy = torch.randn(R * S, L * F, device="cuda")
    pts = (ro[:, None, :] + rd[:, None, :] * t[None, :, None]).reshape(-1, 3).contiguous()
    want = torch.zeros(L, T, F, device="cuda")
    ops.hash_encode_bwd(pts, dy, geom, want, 0, L)
    n_tab, extra = L * T * F, 1024
    reg = PeerRegion(n_tab + extra)
    tab = reg.tensor[:n_tab].view(L, T, F)
    reg.tensor[n_tab:n_tab + extra] = 3.0
    done = torch.zeros(64, dtype=torch.int32, device="cuda")
    chunks = [(0, 2), (2, 8), (8, 9), (9, 14), (14, 16)]
    tiles = ops.hash_bwd_stream_tiles(R * S)
    assert tiles == -(-R * S // 128)
    torch.cuda.synchronize()
    cur, side = torch.cuda.current_stream(), torch.cuda.Stream(priority=-1)
    ev = cur.record_event()
    if form == "rays":
        ops.hash_encode_bwd_rays_stream(ro, rd, t, dy, geom, tab, chunks, done)
    elif form == "rays_lm":                                   # level-major producer: one CTA per 1024 points
        tiles = ops.hash_bwd_lm_ctas(R * S)
        ops.hash_encode_bwd_rays_lm(ro, rd, t, dy.view(R * S, L, F).permute(1, 0, 2).contiguous(), geom, tab, chunks, done)
    else:
        ops.hash_encode_bwd_stream(pts, dy, geom, tab, chunks, done)
    side.wait_event(ev)
    pieces = [(n_tab, extra, 0, 0)] + [(l0 * T * F, (l1 - l0) * T * F, tiles, c) for c, (l0, l1) in enumerate(chunks)]
    with torch.cuda.stream(side):
        reg.all_reduce_stream(pieces, done, scale=0.5, ctas=8)
    cur.wait_stream(side)
    torch.cuda.synchronize()
    assert not reg.timed_out()
    assert done[: len(chunks)].tolist() == [tiles] * len(chunks) and int(done[len(chunks):].abs().sum()) == 0
    assert float((tab - 0.5 * want).norm() / (0.5 * want).norm()) < 1e-6
    assert torch.equal(reg.tensor[n_tab:n_tab + extra], torch.full((extra,), 1.5, device="cuda"))
    with pytest.raises(ValueError):
        ops.hash_encode_bwd_stream(pts, dy, geom, tab, [(0, 4), (5, 16)], done)            # not a partition of the levels
    with pytest.raises(ValueError):
        reg.all_reduce_stream([(2, 8, 0, 0)], None)                                        # offset not a multiple of 4
    reg.close()
