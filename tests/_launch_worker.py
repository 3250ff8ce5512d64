"""Worker for tests/test_cpu.py::test_launch_spawns_world2 (must be importable by the spawned processes)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def worker(outdir, tag):
    import torch

    from faceoff_b200 import distributed as dist

    r, w = dist.get_rank(), dist.get_world_size()
    t = torch.tensor([float(r + 1)])
    dist.all_reduce(t)
    local_ok = dist.get_local_rank() == r      # raises ValueError if the per-machine group was not created
    with open(os.path.join(outdir, f"rank{r}.txt"), "w") as f:
        f.write(f"{tag} {r} {w} {t.item()} {int(local_ok)} {int(dist.is_primary())}")
    dist.synchronize()
