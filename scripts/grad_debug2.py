import sys, torch
sys.path.insert(0, '.')
import torch.nn.functional as F
from oracle import unet_oracle as O
from floodplanet_code_b200 import ops
from floodplanet_code_b200.unet import UNet
from floodplanet_code_b200.loss import MaskedCrossEntropyLoss
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
def rel(a, b): return float((a.double()-b.double()).norm()/b.double().norm().clamp_min(1e-30))
calls = []
orig_apply = ops.bn_relu_bwd_apply
def spy(da, y, dy, scale, shift, coef):
    orig_apply(da, y, dy, scale, shift, coef)
    calls.append((da, y, dy, scale.clone(), shift.clone(), coef.clone()))
ops.bn_relu_bwd_apply = spy
wcalls = []
orig_w = ops.conv3x3_wgrad
def spyw(x, dy, dw, ws, cin_real):
    orig_w(x, dy, dw, ws, cin_real); wcalls.append((x, dy, dw.clone(), cin_real))
ops.conv3x3_wgrad = spyw
n, s = 2, 64
sd = O.init_state_dict(4, 3, seed=0)
m = UNet(4, 3); m.load_state_dict(sd); m = m.cuda().train()
b = O.synthetic_batch(n, 4, s, s, seed=1, block=8, device='cuda')
logits = m(b['image'])
loss = MaskedCrossEntropyLoss(0)(logits, b['target']); loss.backward()
torch.cuda.synchronize()
da, y, dy, scale, shift, coef = calls[0]
gamma = m.up4.conv.double_conv[4].weight.detach(); beta = m.up4.conv.double_conv[4].bias.detach()
yf = y.float().permute(0,3,1,2).contiguous().requires_grad_(True)
out = F.relu(F.batch_norm(yf, None, None, gamma, beta, True, 0.1, 1e-5))
out.backward(da.float().permute(0,3,1,2).contiguous())
ref = yf.grad.permute(0,2,3,1)
print("layer17 dy vs torch BN-bwd on same inputs: rel", rel(dy.float(), ref))
print("  per-channel sum(dy)/sum|dy| ours max", float((dy.float().sum((0,1,2)).abs()/dy.float().abs().sum((0,1,2))).max()),
      " ref max", float((ref.sum((0,1,2)).abs()/ref.abs().sum((0,1,2))).max()))
# forward consistency: scale/shift vs torch stats of the bf16 y
mean = yf.detach().mean((0,2,3)); var = yf.detach().var((0,2,3), unbiased=False)
sc_ref = gamma*torch.rsqrt(var+1e-5); sh_ref = beta-mean*sc_ref
print("  scale rel", rel(scale, sc_ref), "shift rel", rel(shift, sh_ref))
x, dyw, dw, cr = wcalls[0]
refw = torch.nn.grad.conv2d_weight(x.float().permute(0,3,1,2).contiguous(), (dyw.shape[3], x.shape[3], 3, 3), dyw.float().permute(0,3,1,2).contiguous(), padding=1)
print("layer17 wgrad vs torch on same inputs: rel", rel(dw, refw[:, :cr]))
refw2 = torch.nn.grad.conv2d_weight(x.float().permute(0,3,1,2).contiguous(), (dyw.shape[3], x.shape[3], 3, 3), ref.permute(0,3,1,2).contiguous(), padding=1)
print("wgrad(torch dy) vs wgrad(our dy):", rel(refw, refw2))
for i in (1, 2, 5):
    x, dyw, dw, cr = wcalls[i]
    refw = torch.nn.grad.conv2d_weight(x.float().permute(0,3,1,2).contiguous(), (dyw.shape[3], x.shape[3], 3, 3), dyw.float().permute(0,3,1,2).contiguous(), padding=1)
    print(f"wgrad call {i} x{tuple(x.shape)} dy{tuple(dyw.shape)} rel", rel(dw, refw[:, :cr]))
