"""Process launcher with the reference's signature (reference distributed/launch.py:22-92): one process per
GPU, NCCL process group over a localhost TCP rendezvous, per-machine local group."""
import os
import socket

import torch
from torch import distributed as dist
from torch import multiprocessing as mp

from . import distributed as dist_fn


def find_free_port():
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def launch(fn, n_gpu_per_machine, n_machine=1, machine_rank=0, dist_url=None, args=(), backend="nccl"):
    """Same positional signature as the reference; ``backend`` (extra, keyword) exists so the spawn / rendezvous /
    local-group logic can be exercised with gloo on a machine without GPUs."""
    world_size = n_machine * n_gpu_per_machine
    if world_size <= 1:
        fn(*args)
        return
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    if dist_url == "auto":
        if n_machine != 1:
            raise ValueError('dist_url="auto" not supported in multi-machine jobs')
        dist_url = f"tcp://127.0.0.1:{find_free_port()}"
    if n_machine > 1 and dist_url.startswith("file://"):
        raise ValueError("file:// is not a reliable init method in multi-machine jobs. Prefer tcp://")
    mp.spawn(distributed_worker, nprocs=n_gpu_per_machine,
             args=(fn, world_size, n_gpu_per_machine, machine_rank, dist_url, args, backend), daemon=False)


def distributed_worker(local_rank, fn, world_size, n_gpu_per_machine, machine_rank, dist_url, args,
                       backend="nccl"):
    if backend == "nccl" and not torch.cuda.is_available():
        raise OSError("CUDA is not available. Please check your environments")
    global_rank = machine_rank * n_gpu_per_machine + local_rank
    if backend == "nccl":
        if n_gpu_per_machine > torch.cuda.device_count():
            raise ValueError(
                f"specified n_gpu_per_machine larger than available device ({torch.cuda.device_count()})")
        torch.cuda.set_device(local_rank)
    try:
        dist.init_process_group(backend=backend, init_method=dist_url, world_size=world_size, rank=global_rank)
    except Exception as exc:
        raise OSError(f"failed to initialize {backend} groups") from exc
    dist_fn.synchronize()
    if dist_fn.LOCAL_PROCESS_GROUP is not None:
        raise ValueError("faceoff_b200.distributed.LOCAL_PROCESS_GROUP is not None")
    n_machine = world_size // n_gpu_per_machine
    for i in range(n_machine):
        ranks = list(range(i * n_gpu_per_machine, (i + 1) * n_gpu_per_machine))
        pg = dist.new_group(ranks)
        if i == machine_rank:
            dist_fn.LOCAL_PROCESS_GROUP = pg
    fn(*args)
