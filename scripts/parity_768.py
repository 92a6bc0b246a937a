"""Ad-hoc: forward parity at the M3FD shape 768x1024 (BASELINE configs[3]) against the CPU oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import paif_b200
from oracle import fusion_oracle as fo
torch.manual_seed(0)
net = paif_b200.Network_Fusion_Searched(32, None, paif_b200.fusion_at)
sd = {k: v.clone() for k, v in net.state_dict().items()}
net = net.cuda().eval()
g = torch.Generator().manual_seed(5)
ir, vis = torch.rand(1, 1, 768, 1024, generator=g), torch.rand(1, 3, 768, 1024, generator=g)
ref = fo.fusion_forward(sd, paif_b200.fusion_at, ir, vis)
for eng in ("direct", "auto"):
    net.conv_engine = eng
    with torch.no_grad():
        out = net(ir.cuda(), vis.cuda())
    print("768x1024 engine=%s max-abs err %.3e" % (eng, (out.cpu() - ref).abs().max().item()), flush=True)
# odd, non-multiple-of-4 size
ir, vis = torch.rand(2, 1, 77, 203, generator=g), torch.rand(2, 3, 77, 203, generator=g)
ref = fo.fusion_forward(sd, paif_b200.fusion_at, ir, vis)
net.conv_engine = "auto"
with torch.no_grad():
    out = net(ir.cuda(), vis.cuda())
print("77x203 engine=auto max-abs err %.3e" % (out.cpu() - ref).abs().max().item())
# smallest legal size and a wide, short one; forward and backward
for shp in ((1, 10, 10), (3, 11, 300), (1, 130, 10)):
    B, H, W = shp
    ir, vis = torch.rand(B, 1, H, W, generator=g), torch.rand(B, 3, H, W, generator=g)
    cot = torch.randn(B, 1, H, W, generator=g)
    ir_r, vis_r = ir.clone().requires_grad_(True), vis.clone().requires_grad_(True)
    ref = fo.fusion_forward(sd, paif_b200.fusion_at, ir_r, vis_r)
    gi, gv = torch.autograd.grad(ref, [ir_r, vis_r], cot)
    a, v = ir.cuda().requires_grad_(True), vis.cuda().requires_grad_(True)
    out = net(a, v)
    out.backward(cot.cuda())
    rl = ((a.grad.cpu() - gi).norm() / gi.norm()).item()
    print("%s engine=auto fwd err %.3e  grad rel-L2 %.3e" % (shp, (out.detach().cpu() - ref.detach()).abs().max().item(), rl))
