import os, sys, torch
sys.path.insert(0, os.getcwd())
import hybrid_vit_cascade_b200 as hvc
from oracle import vit_oracle as O
c = torch.load("tests/golden/encoder.pt", weights_only=False)["encoder"]
m = hvc.XrayConditioningModule(img_size=64, in_channels=1, embed_dim=64, num_views=2, time_embed_dim=32, cond_dim=96).cuda().train()
m.load_state_dict(c["sd"], strict=True)
xr = c["xrays"].cuda().requires_grad_(True)
ctx, cond, feats = m(xr, c["t"].cuda())
for a, b, n in ((ctx, c["ctx"], "ctx"), (cond, c["cond"], "cond"), (feats, c["feats"], "feats")):
    print(n, "max_rel", O.max_rel(a, b))
sd = m.state_dict()
for k, v in c["sd_after"].items():
    if "num_batches" not in k: print(k, O.max_rel(sd[k], v))
loss = sum((o * r.cuda()).sum() for o, r in zip((ctx, cond, feats), c["r"]))
loss.backward()
for k, p in m.named_parameters():
    print(k, "cos", round(O.cosine(p.grad, c["pgrad"][k].cuda()), 5), "rel", round(O.max_rel(p.grad, c["pgrad"][k]), 4))
print("xgrad cos", O.cosine(xr.grad, c["xgrad"].cuda()))
