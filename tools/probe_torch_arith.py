"""GPU probe: which exact fp32 operation sequences torch's CUDA kernels use for the few ops
whose rounding decides discrete outcomes downstream (run via gpurun; prints a table)."""
import torch
torch.manual_seed(0)
dev = "cuda"
M = 1 << 20
a = torch.randn(M, 3, device=dev); b = torch.randn(M, 3, device=dev) * 5
p = a * b
s = p.sum(-1)
c = {
 "(p0+p1)+p2": (p[:, 0] + p[:, 1]) + p[:, 2],
 "(p0+p2)+p1": (p[:, 0] + p[:, 2]) + p[:, 1],
 "(p1+p2)+p0": (p[:, 1] + p[:, 2]) + p[:, 0],
 "fma chain a0b0,a1b1,a2b2": torch.addcmul(torch.addcmul(p[:, 0], a[:, 1], b[:, 1]), a[:, 2], b[:, 2]),
}
print("== (a*b).sum(-1) over 3 elements")
for k, v in c.items():
    print(f"  {k:28s} bit-equal fraction {(v == s).float().mean().item():.6f}")
s2 = p.sum(-1, keepdim=True)
print("  keepdim same as not:", torch.equal(s2[:, 0], s))
# norm over last dim (3)
g = torch.randn(M, 3, device=dev) * 3
n = g.norm(dim=-1)
q = g * g
c = {
 "sqrt((q0+q1)+q2)": torch.sqrt((q[:, 0] + q[:, 1]) + q[:, 2]),
 "sqrt((q0+q2)+q1)": torch.sqrt((q[:, 0] + q[:, 2]) + q[:, 1]),
 "sqrt(fma(g2,g2,fma(g1,g1,q0)))": torch.sqrt(torch.addcmul(torch.addcmul(q[:, 0], g[:, 1], g[:, 1]), g[:, 2], g[:, 2])),
 "sqrt(fma(g1,g1,fma(g2..)))": torch.sqrt(torch.addcmul(torch.addcmul(q[:, 0], g[:, 2], g[:, 2]), g[:, 1], g[:, 1])),
}
print("== g.norm(dim=-1) over 3 elements")
for k, v in c.items():
    print(f"  {k:34s} bit-equal fraction {(v == n).float().mean().item():.6f}")
# norm over dim=1 of [M,3,4]
g3 = torch.randn(M // 4, 3, 4, device=dev)
n3 = g3.norm(dim=1)
q3 = g3 * g3
c = {
 "sqrt((q0+q1)+q2)": torch.sqrt((q3[:, 0] + q3[:, 1]) + q3[:, 2]),
 "sqrt((q0+q2)+q1)": torch.sqrt((q3[:, 0] + q3[:, 2]) + q3[:, 1]),
 "sqrt(fma(g2,g2,fma(g1,g1,q0)))": torch.sqrt(torch.addcmul(torch.addcmul(q3[:, 0], g3[:, 1], g3[:, 1]), g3[:, 2], g3[:, 2])),
}
print("== g.norm(dim=1) of [M,3,4]")
for k, v in c.items():
    print(f"  {k:34s} bit-equal fraction {(v == n3).float().mean().item():.6f}")
# sum of squares (l2_normalize: torch.sum(x**2, dim=-1))
ss = torch.sum(g ** 2, dim=-1)
print("== torch.sum(x**2,-1): (q0+q1)+q2", ((q[:, 0] + q[:, 1]) + q[:, 2] == ss).float().mean().item(),
      " (q0+q2)+q1", ((q[:, 0] + q[:, 2]) + q[:, 1] == ss).float().mean().item())
# sigmoid
x = torch.randn(M, device=dev) * 8
sg = torch.sigmoid(x)
c = {
 "1/(1+exp(-x))": 1.0 / (1.0 + torch.exp(-x)),
 "exp(x)/(1+exp(x))": torch.exp(x) / (1 + torch.exp(x)),
}
print("== sigmoid")
for k, v in c.items():
    print(f"  {k:22s} bit-equal fraction {(v == sg).float().mean().item():.6f}")
# clamp / relu trivial.  F.linear(x,w,b) vs x@w.T + b
xx = torch.randn(4096, 106, device=dev); w = torch.randn(256, 106, device=dev); bb = torch.randn(256, device=dev)
y = torch.nn.functional.linear(xx, w, bb)
y64 = (xx.double() @ w.double().T + bb.double())
print("== F.linear fp32 vs fp64: max rel err", ((y.double() - y64).abs().max() / y64.abs().max()).item(),
      " allow_tf32 =", torch.backends.cuda.matmul.allow_tf32)
