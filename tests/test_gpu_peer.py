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
