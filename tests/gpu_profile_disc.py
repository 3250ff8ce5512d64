import sys, os, json, types
sys.path.insert(0, '/root/repo')
import torch, time
from faceoff_b200.mocoganhd import content_disc, losses, video_disc
dev='cuda'
torch.manual_seed(0)
d3 = video_disc.ModelD_3d(3, "instance", 2, 1e-4, False, 12).to(dev).train()
crit = losses.Relativistic_Average_LSGAN()
x = (torch.rand(1, 6, 11, 256, 256)*2-1).to(dev)
from faceoff_b200.mocoganhd import layers
conv = d3.netD.scale1_layer3[0]
for name, shp in (("L4 3d 256->512 s1", (1,256,3,33,33)), ("L2 3d 64->128 s2", (1,64,6,129,129))):
    m = layers.Conv3d(shp[1], shp[1]*2, 4, stride=1 if "s1" in name else 2, padding=2).cuda()
    xx = torch.randn(shp, device=dev, requires_grad=True)
    y = m(xx); go = torch.randn_like(y)
    for _ in range(2): y = m(xx); y.backward(go)
    torch.cuda.synchronize()
    a,b,c,d = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    a.record(); y = m(xx); b.record(); y.backward(go); c.record(); torch.cuda.synchronize()
    fl = 2.0*m.weight.numel()*y[0,0].numel()
    print(name, "fwd %.2f ms %.1f TF/s; bwd %.2f ms %.1f TF/s" % (a.elapsed_time(b), fl/a.elapsed_time(b)/1e9, b.elapsed_time(c), 2*fl/b.elapsed_time(c)/1e9))
