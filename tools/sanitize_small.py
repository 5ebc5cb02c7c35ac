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
