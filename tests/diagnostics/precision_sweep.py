"""Diagnostic (GPU, run by hand: `python tests/diagnostics/precision_sweep.py` from the repo root; lives under
tests/ because it uses the oracle, which only test code may import): distance of the CUDA path from (a) the fp32 oracle and (b) the oracle with
bf16 storage emulated, across problem sizes.  (b) isolates kernel bugs from bf16 rounding."""
import copy, sys, torch
sys.path.insert(0, '.')
from oracle import unet_oracle as O
from floodplanet_code_b200.unet import UNet
from floodplanet_code_b200.loss import MaskedCrossEntropyLoss
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
def rel(a, b): return float((a.double()-b.double()).norm()/b.double().norm().clamp_min(1e-30))
for (n, s) in [(2, 32), (2, 64), (2, 128), (4, 256), (8, 512)]:
    sd = O.init_state_dict(4, 3, seed=0)
    m = UNet(4, 3); m.load_state_dict(sd); m = m.cuda().train()
    b = O.synthetic_batch(n, 4, s, s, seed=1, block=8, device='cuda')
    sdg = {k: v.cuda() for k, v in sd.items()}
    sd1 = copy.deepcopy(sdg); sd2 = copy.deepcopy(sdg)
    logits = m(b['image'])
    loss = MaskedCrossEntropyLoss(0)(logits, b['target']); loss.backward()
    oloss, opred, ologits, ograds = O.training_step(sd1, b, 0, early_fusion=False)
    with torch.no_grad(): elog = O.unet_forward_bf16_emulated(sd2, b['image'], True)
    named = dict(m.named_parameters())
    errs = sorted(((rel(named[k].grad, g), k) for k, g in ograds.items() if not (k.endswith('.0.bias') or k.endswith('.3.bias'))), reverse=True)
    print(f"n={n} s={s}: logits vs fp32 {rel(logits, ologits):.4f}  vs bf16-emulated {rel(logits, elog):.5f}  emulated vs fp32 {rel(elog, ologits):.4f}  loss rel {abs(float(loss)-float(oloss))/abs(float(oloss)):.5f}  worst grads {[(round(e,4),k) for e,k in errs[:3]]}  median grad err {errs[len(errs)//2][0]:.4f}", flush=True)
    del m, logits, loss
    torch.cuda.empty_cache()
