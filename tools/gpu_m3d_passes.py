import sys; sys.path.insert(0, '.')
import torch, numpy as np
from dspfun_b200 import Plan, REDFT10, REDFT01
D,H,W = 256,1080,1920
x = torch.rand((D,H,W), device='cuda', dtype=torch.float32)
for kind, name in ((REDFT10,'fwd'),(REDFT01,'inv')):
    p = Plan('f',[D,H,W],[kind]*3).profile(True)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3): p.execute_dev(x.data_ptr(), x.data_ptr(), st)
    torch.cuda.synchronize()
    p.pass_stats()
    for _ in range(3): p.execute_dev(x.data_ptr(), x.data_ptr(), st); x.mul_(1e-9)
    torch.cuda.synchronize()
    for s in p.pass_stats(): print(name, s['kernel'], 'axis', s['axis'], 'n', s['n'], 'grid', s['grid'], 'smem', s['smem_bytes'], '%.3f ms' % (s['ms_total']/max(1,s['launches'])), '%.0f GB/s' % (8*s['samples']/(s['ms_total']/max(1,s['launches'])*1e-3)/1e9))
    p.destroy()
