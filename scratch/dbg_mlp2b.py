import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ffb200 import native as nv
from ffb200.models.FactorFields import MLPMixer
torch.manual_seed(3)
mm = MLPMixer(18, 32, num_layers=2, hidden_dim=64).cuda()
x = torch.randn(20000, 18, device='cuda', requires_grad=True)
G = torch.randn(20000, 32, device='cuda')
outs = []
for fused in (1, 0):
    nv.lib().ffb_set_fused_mlp(fused)
    y = mm(x)
    outs.append([y.detach()] + [g.detach() for g in torch.autograd.grad((y * G).sum(), [x] + list(mm.parameters()))])
names = ['y', 'gx', 'gW1', 'gb1', 'gW2']
for nm, a, b in zip(names, *outs):
    print(nm, float((a - b).abs().max() / b.abs().max()))
a, b = outs[0][1], outs[1][1]
e = (a - b).abs().max(1).values
bad = (e > 1e-4 * float(b.abs().max())).nonzero().flatten()
print('rows with gx diff:', len(bad), bad[:10].tolist())
W1, b1 = mm.backbone[0].weight.double(), mm.backbone[0].bias.double()
z = x.detach().double() @ W1.T + b1
for r in bad[:10].tolist():
    print(r, 'min |z| in row', float(z[r].abs().min()), 'err', float(e[r]))
