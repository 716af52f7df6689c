import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lcrnet_b200 import pair_ops as P
os.environ['LCR_ATTN'] = 'tc'
nq, nk = 128, 64
q = torch.zeros(nq, 128).cuda()
k = torch.zeros(nk, 128).cuda()
qo = torch.tensor([0, nq]).cuda(); ko = torch.tensor([0, nk]).cuda()
for name, v in (('key index', torch.arange(nk).float()[:, None].repeat(1, 128)), ('dim index', torch.arange(128).float()[None].repeat(nk, 1)),
                ('one-hot key 5', torch.zeros(nk, 128).index_fill_(0, torch.tensor([5]), 64.0)),
                ('one-hot key 40', torch.zeros(nk, 128).index_fill_(0, torch.tensor([40]), 64.0))):
    out = P.attention(q, k, v.cuda().contiguous(), qo, ko, 1, nq, heads=4).cpu()
    print(name, 'row0[:8]', out[0, :8].tolist(), 'row77[:4]', out[77, :4].tolist(), 'row0[32:36]', out[0, 32:36].tolist())
