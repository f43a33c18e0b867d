"""Numerics study for a tensor-core implementation of the 1e-5 (fp32-parity) mode -- CPU only, TEST INFRASTRUCTURE.

Every nn.Linear with in_features >= 128 of the oracle is replaced, forward AND backward, by a split-operand product
accumulated in fp32, emulating what tcgen05 would compute:
  f16x2   a ~ hi + lo in fp16, 3 MMAs (hi.hi + hi.lo + lo.hi)          4 B / element in shared memory
  bf16x2  a ~ h + m in bf16, 3 MMAs                                     4 B / element
  tf32x3  a ~ hi + lo in tf32 (hardware truncation), 3 kind::tf32 MMAs  8 B / element
  bf16x3  a ~ h + m + l in bf16, 6 MMAs (all terms of order <= 2)        6 B / element
and fields / per-tensor gradients are compared with the fp64 oracle next to the plain fp32 oracle (= the reference).

    python tests/tools/split_precision_study.py 2 256      # result:
      f16x2 : fields 1.6e-6, worst gradient tensor 1.1e-3   (fp16 range: gradient tiles underflow their low halves; FAILS)
      bf16x2: fields 1.7e-5, worst gradient tensor 5.6e-4   (FAILS the 1e-5 mode)
      tf32x3: fields 1.3e-6, worst gradient tensor 5.9e-6   (0.35x the reference's own fp32 gradient error)
      bf16x3: fields 3.1e-7, worst gradient tensor 7.5e-6   (0.49x)
"""
import sys, math, os
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import torch, numpy as np
import torch.nn.functional as F
from oracle import pdg_oracle as O
import pdg_helpers as H

def split(v):
    hi = v.half().float()
    lo = (v - hi).half().float()
    return hi, lo

KIND = {"k": "f16x2"}
def tf32(v):
    return (v.contiguous().view(torch.int32) & ~0x1fff).view(torch.float32)
def split_bf3(v):
    h = v.bfloat16().float(); m = (v - h).bfloat16().float(); l = (v - h - m).bfloat16().float(); return h, m, l
def mm3(a, b):
    k = KIND["k"]
    if k == "f16x2":
        ah, al = split(a); bh, bl = split(b)
        return ah @ bh + (ah @ bl + al @ bh)
    if k == "tf32x3":
        ah = tf32(a); al = tf32(a - ah); bh = tf32(b); bl = tf32(b - bh)
        return ah @ bh + (ah @ bl + al @ bh)
    ah, am, al = split_bf3(a); bh, bm, bl = split_bf3(b)
    if k == "bf16x2":
        return ah @ bh + (ah @ bm + am @ bh)
    return ah @ bh + (ah @ bm + am @ bh) + (am @ bm + ah @ bl + al @ bh)

class SplitLinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b):
        ctx.save_for_backward(x, w)
        return mm3(x, w.t()) + b
    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        return mm3(dy, w), mm3(dy.t(), x), dy.sum(0)

MODE = {"on": False}
orig_linear = F.linear
def patched(x, w, b=None):
    if MODE["on"] and x.dtype == torch.float32 and w.shape[1] >= 128:
        return SplitLinear.apply(x, w, b)
    return orig_linear(x, w, b)
O.F.linear = patched

def grads(sd, batch, stats, dtype, scale=1.0, split_on=False):
    from collections import OrderedDict
    MODE["on"] = split_on
    p = OrderedDict((k, v.detach().clone().to(dtype).requires_grad_(True)) for k, v in sd.items())
    total, nmse, div, pred = O.train_loss(p, batch, stats, 10, True, 10.0, dtype)
    g = torch.autograd.grad(total * scale, list(p.values()), allow_unused=True)
    MODE["on"] = False
    return {k: (gi / scale if gi is not None else torch.zeros_like(v)) for k, gi, v in zip(p.keys(), g, p.values())}, pred.detach()

n_graphs, nodes = int(sys.argv[1]), int(sys.argv[2])
samples, graphs, batch, stats = H.synthetic_batch(n_graphs, nodes, seed0=69)
sd = O.init_state_dict(seed=69)
g64, p64 = grads(sd, batch, stats, torch.float64)
g32, p32 = grads(sd, batch, stats, torch.float32)
# scale so that max |dL/dpred| = 2^8
pred = p32.clone().requires_grad_(True)
print("pred err fp32", H.rel_err(p32, p64))
for kind in ("f16x2", "bf16x2", "tf32x3", "bf16x3"):
    KIND["k"] = kind; S = 1.0
    gs, ps = grads(sd, batch, stats, torch.float32, scale=S, split_on=True)
    worst = (0, 0, ""); worst32 = (0, 0, "")
    ratios = []
    for k in O.STATE_KEYS:
        e = H.rel_err(gs[k], g64[k]); r = H.rel_err(g32[k], g64[k])
        ratios.append(max(e[0] / max(r[0], 1e-5), e[1] / max(r[1], 1e-5)))
        if e[0] > worst[0]: worst = (e[0], e[1], k)
        if r[0] > worst32[0]: worst32 = (r[0], r[1], k)
    print(f"{kind}: pred err {H.rel_err(ps, p64)}, worst split grad err {worst}, worst fp32 {worst32}, max ratio split/fp32 {max(ratios):.2f}")
