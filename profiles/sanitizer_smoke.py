import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
from cxl_speckv_b200 import codec
from cxl_speckv_b200.tier import HostTier
rng=np.random.default_rng(0)
for G,n in ((131072,3),(2048,40),(32768,5),(8192,9)):
    x=rng.standard_normal(n*G).astype(np.float16)
    x[G//2:G//2+700]=0.5
    x[G + G//3: 2*G]=0.0                       # a partially filled block: zero tail (zero-region shortcut, split decode)
    xd=torch.from_numpy(x).cuda()
    c=codec.compress(xd,G); y=codec.decompress(c); torch.cuda.synchronize()
    idx=torch.tensor([1,0],dtype=torch.int32,device='cuda'); z=codec.decompress_indexed(c,idx); torch.cuda.synchronize()
# packed emission of the tier's page groups (look-back over CTA status words), mixed page kinds
G,n=2048,600
x=rng.standard_normal(n*G).astype(np.float16).reshape(n,G)
x[5:40]=0.0; x[50:60]=0.25; x[70,3]=np.inf; x[80:90,G//2:]=0.0
xd=torch.from_numpy(x.reshape(-1)).cuda()
t=HostTier(pool_bytes=8<<20)
ids=np.arange(n,dtype=np.uint64)
t.offload(xd,G,ids); y=t.restore(ids,G,torch.float16); torch.cuda.synchronize()
want=codec.decompress(codec.compress(xd,G))
assert torch.equal(y.view(torch.int16),want.view(torch.int16))
t.close()
print("done")
