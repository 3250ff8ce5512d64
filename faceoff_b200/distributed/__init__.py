"""Data-parallel helpers with the reference's ``distributed`` package API (reference distributed/__init__.py:1-13):
get_rank, get_local_rank, is_primary, synchronize, get_world_size, all_reduce, all_gather, reduce_dict,
data_sampler, LOCAL_PROCESS_GROUP, launch.  One process per GPU over torch.distributed (NCCL on NVLink /
NVSwitch; gloo for CPU tests).  The per-step collectives of the training path are fused by
faceoff_b200.parallel.FusedDataParallel into one bucketed all-reduce.
"""
from .distributed import (  # noqa: F401
    LOCAL_PROCESS_GROUP,
    all_gather,
    all_reduce,
    data_sampler,
    get_local_rank,
    get_rank,
    get_world_size,
    is_primary,
    reduce_dict,
    synchronize,
)
from .launch import launch  # noqa: F401
