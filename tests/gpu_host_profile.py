"""Host-side (Python + C ABI) enqueue cost of one training step: cProfile over a few steps at a small batch.

    python tests/gpu_host_profile.py [clips] [steps]
"""
import cProfile
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from faceoff_b200.losses import mse_loss  # noqa: E402
from faceoff_b200.vqvae import VQVAE  # noqa: E402
from oracle import faceoff_oracle as O  # noqa: E402  (seeded weights only)


def main():
    clips = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    dev = torch.device("cuda", 0)
    model = VQVAE(in_channel=6)
    model.load_state_dict(O.init_vqvae_params(seed=0))
    model = model.to(dev).train()
    img = torch.empty(clips * 30, 6, 256, 256, device=dev).uniform_(-1, 1)
    gt = torch.empty(clips * 30, 3, 256, 256, device=dev).uniform_(-1, 1)

    def step():
        model.zero_grad(set_to_none=True)
        out, latent = model.forward_with_ids(img, clips)[:2]
        loss = mse_loss(out, gt) + latent.mean()
        loss.backward()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    host = (time.perf_counter() - t0) / steps * 1e3
    torch.cuda.synchronize()
    total = (time.perf_counter() - t0) / steps * 1e3
    print(f"clips {clips}: host enqueue {host:.2f} ms/step, wall {total:.2f} ms/step")
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(steps):
        step()
    pr.disable()
    torch.cuda.synchronize()
    st = pstats.Stats(pr)
    st.sort_stats("tottime").print_stats(22)


if __name__ == "__main__":
    main()
