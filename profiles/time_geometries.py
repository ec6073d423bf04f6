import sys; sys.path.insert(0,'/root/repo')
import torch, time
# 1 GiB of fp16 per geometry: kv GB/s = uncompressed bytes / kernel time (fast + flagged pass)
from cxl_speckv_b200 import codec
for G,n in ((2048,262144),(8192,65536),(16384,32768),(32768,16384),(65536,8192),(131072,4096),(262144,2048)):
    x=torch.randn(n*G,device='cuda').half()
    c=codec.compress(x,G); y=codec.decompress(c); torch.cuda.synchronize()
    for name,fn in (('compress',lambda: codec.compress(x,G,out=c)),('decompress',lambda: codec.decompress(c,out=y))):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20): fn()
        b.record(); torch.cuda.synchronize()
        ms=a.elapsed_time(b)/20
        print(G,n,name,round(ms*1e3,1),'us', round(n*G*2/ms/1e6,1),'GB/s kv')
