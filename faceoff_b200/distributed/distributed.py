"""Collective helpers; semantics follow reference distributed/distributed.py:12-143 (every helper degrades
to a no-op when no process group exists, all_reduce is an in-place SUM, all_gather moves picklable objects)."""
import pickle

import torch
from torch import distributed as dist
from torch.utils import data

LOCAL_PROCESS_GROUP = None


def _active() -> bool:
    return dist.is_available() and dist.is_initialized()


def get_rank():
    return dist.get_rank() if _active() else 0


def is_primary():
    return get_rank() == 0


def get_local_rank():
    if not _active():
        return 0
    if LOCAL_PROCESS_GROUP is None:
        raise ValueError("faceoff_b200.distributed.LOCAL_PROCESS_GROUP is None")
    return dist.get_rank(group=LOCAL_PROCESS_GROUP)


def get_world_size():
    return dist.get_world_size() if _active() else 1


def synchronize():
    if _active() and dist.get_world_size() > 1:
        dist.barrier()


def all_reduce(tensor, op=dist.ReduceOp.SUM):
    """In-place reduction over all ranks (SUM by default, reference :64-72); returns the tensor."""
    if get_world_size() > 1:
        dist.all_reduce(tensor, op=op)
    return tensor


def _comm_device():
    backend = dist.get_backend()
    return torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")


def all_gather(data):
    """Gather one picklable object per rank into a list (reference :75-107)."""
    world = get_world_size()
    if world == 1:
        return [data]
    dev = _comm_device()
    payload = torch.frombuffer(bytearray(pickle.dumps(data)), dtype=torch.uint8).to(dev)
    sizes = torch.zeros(world, dtype=torch.int64, device=dev)
    sizes[get_rank()] = payload.numel()
    dist.all_reduce(sizes)
    sizes = sizes.tolist()
    cap = max(sizes)
    buf = torch.zeros(cap, dtype=torch.uint8, device=dev)
    buf[: payload.numel()] = payload
    parts = [torch.empty(cap, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(parts, buf)
    return [pickle.loads(p[:n].cpu().numpy().tobytes()) for p, n in zip(parts, sizes)]


def reduce_dict(input_dict, average=True):
    """Reduce a dict of same-shaped tensors to rank 0 (reference :110-132)."""
    world = get_world_size()
    if world < 2:
        return input_dict
    with torch.no_grad():
        keys = sorted(input_dict.keys())
        stacked = torch.stack([input_dict[k] for k in keys], 0)
        dist.reduce(stacked, dst=0)
        if dist.get_rank() == 0 and average:
            stacked /= world
        return dict(zip(keys, stacked))


def data_sampler(dataset, shuffle, distributed):
    if distributed:
        return data.distributed.DistributedSampler(dataset, shuffle=shuffle)
    return data.RandomSampler(dataset) if shuffle else data.SequentialSampler(dataset)
