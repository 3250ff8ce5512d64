"""How the length of a tensor-core accumulation chain shows in a weight gradient (bench.py dp_check).

    python tests/gpu_accum_bias.py

Conv3d 128 -> 128 weight gradient (wgrad_igemm, 16 splits) on bf16-representable random activations of `clips` clips of
30 x 64 x 64 positions: the centre tap dW[:, :, 1, 1, 1] = dy^T x is compared with an fp64 GEMM of the same operands
 (a) from ONE launch over all clips           (chain = clips x 480 MMAs of K = 16 per TMEM accumulator),
 (b) as the fp32 sum of per-clip launches      (chain = 480, what N data-parallel ranks compute before the all-reduce).
Printed per variant: max-normalised error, relative error of the Frobenius norm (signed: negative = shrink).
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from faceoff_b200 import ops  # noqa: E402
from faceoff_b200.ops import FORM_S1  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    for clips in (1, 2, 8, 32):
        x = torch.randn(clips, 30, 64, 64, 128, device=dev).to(torch.bfloat16)
        dy = (torch.randn(clips, 30, 64, 64, 128, device=dev) * 1e-3).to(torch.bfloat16)
        ref = torch.zeros(128, 128, dtype=torch.float64, device=dev)
        for c in range(clips):
            ref += dy[c].reshape(-1, 128).double().t() @ x[c].reshape(-1, 128).double()
        one = torch.empty(128, 128, 3, 3, 3, dtype=torch.float32, device=dev)
        ops.wgrad(FORM_S1, 3, 3, (dy, 128, 0), (x, 128, 0), one, m_axis=0)
        per = torch.zeros(128, 128, 3, 3, 3, dtype=torch.float32, device=dev)
        tmp = torch.empty_like(per)
        for c in range(clips):
            ops.wgrad(FORM_S1, 3, 3, (dy[c:c + 1], 128, 0), (x[c:c + 1], 128, 0), tmp, m_axis=0)
            per += tmp
        torch.cuda.synchronize()
        for name, w in (("one launch", one), ("per-clip launches, fp32 sum", per)):
            a = w[:, :, 1, 1, 1].double()
            e_max = ((a - ref).abs().max() / ref.abs().max()).item()
            e_norm = ((a.norm() - ref.norm()) / ref.norm()).item()
            print(f"clips {clips:2d} {name:28s}: max-normalised err {e_max:.3e}  norm rel err {e_norm:+.3e}")


if __name__ == "__main__":
    main()
