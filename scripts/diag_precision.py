"""GPU diagnostic: absolute error of the projection's first layer z = x W1^T + b1 (pre-ReLU sign matters for the
backward gate) for each precision mode, against float64."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from two_towers_overlords_b200 import ops, _native as N

torch.manual_seed(0)
M, H, P = 4096, 384, 512
x = torch.nn.functional.normalize(torch.randn(M, H), dim=1)
l1, l2 = torch.nn.Linear(H, P), torch.nn.Linear(P, P)
z64 = (x.double() @ l1.weight.double().T + l1.bias.double())
h64 = z64.clamp_min(0)
y64 = h64 @ l2.weight.double().T + l2.bias.double()
zt = torch.nn.functional.linear(x, l1.weight, l1.bias)
print("torch cpu fp32 : max|dz| %.3e rms %.3e" % ((zt - z64).abs().max(), (zt - z64).pow(2).mean().sqrt()))
dev = "cuda"
lib = N.load()
for prec in ("fp32", "bf16x3", "bf16"):
    pc = N.PRECISIONS[prec]
    xd = x.to(dev); W1, b1, W2, b2 = (t.detach().to(dev).contiguous() for t in (l1.weight, l1.bias, l2.weight, l2.bias))
    h = torch.empty(M, P, device=dev); y = torch.empty(M, P, device=dev)
    wsb = lib.tt_mlp_ws_bytes(M, H, P, pc); ws = N.workspace(wsb, dev)
    N.check(lib.tt_encode_fwd(N.ptr(xd), M, H, P, N.ptr(W1), N.ptr(b1), N.ptr(W2), N.ptr(b2), N.ptr(h), N.ptr(y), pc,
                              N.ptr(ws), wsb, N.stream()), "fwd")
    torch.cuda.synchronize()
    hh = h.cpu().double()
    pos = z64 > 1e-3
    dz = (hh - h64)[pos]
    flips = int(((hh > 0) != (z64 > 0)).sum())
    dy = (y.cpu().double() - y64)
    print(f"{prec:7s}: h abs err max {dz.abs().max():.3e} rms {dz.pow(2).mean().sqrt():.3e} | gate flips {flips} of {M*P} "
          f"| y rel {dy.norm()/y64.norm():.3e}")
