"""One rank of a multi-GPU run of the UNMODIFIED trainer (one process per GPU over NCCL / NVLink):

    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29500 \\
        /path/to/this/repo/launch_rank.py /path/to/Human-Body-Reconstruction/train_hash2.py --hash_size 19 ...

Each process pins its own GPU with torch.cuda.set_device(LOCAL_RANK) (the other devices stay visible so the ranks can map
each other's gradient buffers; nn.DataParallel at train_hash2.py:127 is told to stay on that one device), joins the NCCL
group, and sets HBR_AUTO_DP=1: the drop-in Volume_Renderer then, at its first vol_render call, broadcasts rank 0's
encoder tables and MLP parameters to every rank (the script does not seed, so each process initialised its own -- and
keeps its own RNG stream, so the DataLoader of every rank shuffles the rays differently) and attaches the gradient
all-reduce (dist.attach_grad_allreduce) that runs inside loss.backward().  Each rank trains on its own num_batch rays
per step: the global batch is WORLD_SIZE x --num_batch.
"""
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    if len(sys.argv) < 2:
        print(__doc__)
        sys.exit(2)
    import torch
    from human_body_reconstruction_b200 import dist as hdist
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if torch.cuda.is_available():
        torch.cuda.set_device(local)
    hdist.init_from_env()
    os.environ["HBR_AUTO_DP"] = "1"
    # nn.DataParallel(module) with device_ids=None would span every visible GPU inside this process: keep it on ours
    _dp_init = torch.nn.DataParallel.__init__

    def _one_device(self, module, device_ids=None, output_device=None, dim=0):
        if device_ids is None and torch.cuda.is_available():
            device_ids = [torch.cuda.current_device()]
        _dp_init(self, module, device_ids=device_ids, output_device=output_device, dim=dim)

    torch.nn.DataParallel.__init__ = _one_device
    import launch
    launch.run(sys.argv[1], sys.argv[2:])


if __name__ == "__main__":
    main()
