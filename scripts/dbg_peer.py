"""torchrun --nproc-per-node N scripts/dbg_peer.py : correctness + timing of the peer-memory all-reduce against NCCL."""
import os, sys, json, signal, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from human_body_reconstruction_b200 import dist as hdist
from human_body_reconstruction_b200.peer import PeerRegion

rank, world = hdist.init_from_env("nccl")
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
out = {"world": world}
n = 16 * 2 ** 19 * 2 + 14228

def bench(fn, iters=10):
    for _ in range(3): fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)

def check_region(reg, tag):
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    src = torch.randn(reg.n, device=dev, generator=g)
    ref = src.clone()
    dist.all_reduce(ref, op=dist.ReduceOp.SUM)
    ref /= world
    reg.tensor.copy_(src)
    reg.all_reduce(scale=1.0 / world)
    torch.cuda.synchronize()
    err = float((reg.tensor - ref).abs().max())
    # bit-identical across ranks?
    mine = reg.tensor.clone()
    other = mine.clone()
    dist.broadcast(other, src=0)
    same = bool(torch.equal(mine, other))
    out[tag + "_max_abs_err_vs_nccl"] = err
    out[tag + "_identical_across_ranks"] = same
    out[tag + "_timed_out"] = reg.timed_out()
    for ctas in (32, 64, 128):
        out[f"{tag}_ms_ctas{ctas}"] = bench(lambda: reg.all_reduce(scale=1.0 / world, ctas=ctas))
    out[tag + "_timed_out_after"] = reg.timed_out()

try:
    reg = PeerRegion(n, transport="ipc")
    check_region(reg, "ipc")
except Exception as e:
    out["ipc_error"] = repr(e)[:300]

x = torch.randn(n, device=dev)
out["nccl_avg_ms"] = bench(lambda: dist.all_reduce(x, op=dist.ReduceOp.AVG))
out["GBps_algo_nccl"] = n * 4 / out["nccl_avg_ms"] / 1e6

def alarm(*a): raise TimeoutError("symm rendezvous timed out")
signal.signal(signal.SIGALRM, alarm)
if os.environ.get("HBR_TRY_SYMM", "1") == "1":
    try:
        signal.alarm(60)
        reg2 = PeerRegion(n, transport="symm")
        signal.alarm(0)
        out["symm_multicast_ptr"] = hex(reg2.multicast_ptr)
        check_region(reg2, "symm")
    except BaseException as e:
        signal.alarm(0)
        out["symm_error"] = repr(e)[:400]
if rank == 0:
    print(json.dumps(out), flush=True)
sys.stdout.flush()
torch.cuda.synchronize()
os._exit(0)
