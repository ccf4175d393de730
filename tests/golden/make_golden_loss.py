"""Golden fixtures for DirectRegressionLoss from the REAL reference (authoring container only):
    python tests/golden/make_golden_loss.py  ->  tests/golden/direct_loss.pt"""
import sys, os, torch
sys.path.insert(0, "/root/reference"); sys.path.insert(0, "/root/reference/direct_regression")
from model_direct import DirectRegressionLoss, compute_ssim_loss
g = torch.Generator().manual_seed(77)
out = {}
for name, shape in (("small", (2, 1, 16, 20, 24)), ("cube32", (1, 1, 32, 32, 32))):
    pred = (torch.rand(shape, generator=g) * 2 - 1).requires_grad_(True)
    target = torch.rand(shape, generator=g) * 2 - 1
    target = (0.7 * target + 0.3 * pred.detach())            # correlated, like a half-trained prediction
    res = DirectRegressionLoss(1.0, 0.5)(pred, target)
    res["total_loss"].backward()
    out[name] = dict(pred=pred.detach(), target=target, total=res["total_loss"].detach(), l1=res["l1_loss"].detach(),
                     ssim=res["ssim_loss"].detach(), dpred=pred.grad.clone())
torch.save(out, os.path.join(os.path.dirname(os.path.abspath(__file__)), "direct_loss.pt"))
print({k: float(v["total"]) for k, v in out.items()})
