import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
from cxl_speckv_b200 import codec
rng=np.random.default_rng(0)
for G,n in ((131072,3),(2048,40),(32768,5),(8192,9)):
    x=rng.standard_normal(n*G).astype(np.float16)
    x[G//2:G//2+700]=0.5
    xd=torch.from_numpy(x).cuda()
    c=codec.compress(xd,G); y=codec.decompress(c); torch.cuda.synchronize()
    idx=torch.tensor([1,0],dtype=torch.int32,device='cuda'); z=codec.decompress_indexed(c,idx); torch.cuda.synchronize()
print("done")
