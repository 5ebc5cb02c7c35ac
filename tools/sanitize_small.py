import torch, io, contextlib, sys
sys.path.insert(0, '.')
from xumx_slicq_b200 import NSGTBase, make_filterbanks
dev = torch.device('cuda:0')
with contextlib.redirect_stdout(io.StringIO()):
    base = NSGTBase('bark', 262, 32.9, device=dev)
nsgt, insgt = make_filterbanks(base)
x = torch.rand(1, 2, 40000, device=dev) * 2 - 1
X, Xm = nsgt.forward_with_norm(x)
y = insgt(X, 40000)
masks = [torch.rand((2,) + tuple(Xb.shape[:-1]), device=dev) for Xb in X]
ym = insgt.forward_masked(X, masks, 40000)
torch.cuda.synchronize()
print('err', float((y - x).abs().max()), ym.shape)
# round 2: gradients in both directions (adjoint plans), the generic slice kernels and the mirrored-bin pass (tiny-mel)
xg = x.clone().requires_grad_(True)
Xg = nsgt(xg)
(0.5 * (insgt(Xg, 40000) - 0.3) ** 2).sum().backward()
with contextlib.redirect_stdout(io.StringIO()):
    mel = NSGTBase('mel', 32, 115.5, device=dev)
n2, i2 = make_filterbanks(mel)
xm = torch.rand(1, 2, 9000, device=dev) * 2 - 1
ym2 = i2(n2(xm), 9000)
torch.cuda.synchronize()
print('grad', float(xg.grad.abs().max()), 'tiny-mel err', float((ym2 - xm).abs().max()))
