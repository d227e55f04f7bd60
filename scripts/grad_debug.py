import copy, sys, torch
sys.path.insert(0, '.')
from oracle import unet_oracle as O
from floodplanet_code_b200.unet import UNet
from floodplanet_code_b200.loss import MaskedCrossEntropyLoss
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
def rel(a, b): return float((a.double()-b.double()).norm()/b.double().norm().clamp_min(1e-30))
n, s = 2, 64
sd = O.init_state_dict(4, 3, seed=0)
m = UNet(4, 3); m.load_state_dict(sd); m = m.cuda().train()
b = O.synthetic_batch(n, 4, s, s, seed=1, block=8, device='cuda')
sd1 = {k: v.cuda() for k, v in sd.items()}
logits = m(b['image'])
lf = MaskedCrossEntropyLoss(0)
loss = lf(logits, b['target']); 
dl_ours = torch.autograd.grad(loss, logits, retain_graph=True)[0]
loss.backward()
keys = O.trainable_keys(sd1)
for k in keys: sd1[k].requires_grad_(True)
ologits = O.unet_forward(sd1, b['image'], True)
ologits.retain_grad()
oloss, _ = O.masked_ce(ologits, b['target'], 0)
oloss.backward()
print("dlogits rel", rel(dl_ours, ologits.grad), "norm ratio", float(dl_ours.norm()/ologits.grad.norm()))
named = dict(m.named_parameters())
for k in keys:
    g, og = named[k].grad.double().flatten(), sd1[k].grad.double().flatten()
    cos = float((g @ og) / (g.norm() * og.norm()).clamp_min(1e-30))
    print(f"{k:50s} rel {rel(g, og):8.4f} cos {cos:8.4f} ratio {float(g.norm()/og.norm().clamp_min(1e-30)):8.4f} |og| {float(og.norm()):.3e}")
