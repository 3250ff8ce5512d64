import os, sys, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from faceoff_b200.parallel import FusedDataParallel
from faceoff_b200.vqvae import VQVAE
from faceoff_b200 import vqvae as V
from oracle import faceoff_oracle as O
rank = int(os.environ["RANK"]); dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
dist.init_process_group("gloo")
p = O.init_vqvae_params(seed=0)
img, gt = O.synthetic_clip(2, 4, 64, 64, seed=77)
m = VQVAE(in_channel=6); m.load_state_dict(p); m = m.to(dev).train()
ddp = FusedDataParallel(m)
orig = V.Quantize.apply_ema
def dbg(self, c, s):
    print(rank, "apply_ema called counts sum", c.sum().item(), flush=True)
    orig(self, c, s)
V.Quantize.apply_ema = dbg
e0 = m.quantize_t.embed.clone()
with torch.no_grad():
    out = ddp(img[:4].to(dev))
torch.cuda.synchronize()
print(rank, "training", m.training, m.quantize_t.training, "changed", not torch.equal(e0, m.quantize_t.embed), (e0 - m.quantize_t.embed).abs().max().item(), flush=True)
dist.barrier(); dist.destroy_process_group()
