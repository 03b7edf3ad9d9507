import sys; sys.path.insert(0,'/root/repo')
import torch, torch.nn.functional as F
from musketeer_b200 import ops
g=torch.Generator().manual_seed(0)
for (N,C,H,W,relu,res) in [(1,256,96,96,False,False),(1,64,96,96,True,False),(4,256,96,96,True,True),(1,1024,24,24,True,True)]:
    x=(torch.randn(N,C,H,W,generator=g)*1.5+0.3).cuda().bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_()
    r=torch.randn(N,C,H,W,generator=g).cuda().bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_() if res else None
    gm=(1+0.2*torch.randn(C,generator=g)).cuda().bfloat16().requires_grad_(); bt=(0.2*torch.randn(C,generator=g)).cuda().bfloat16().requires_grad_()
    rm=torch.zeros(C).cuda().bfloat16(); rv=torch.ones(C).cuda().bfloat16()
    y=ops.batch_norm(x,gm,bt,rm,rv,residual=r,relu=relu,training=True)
    dy=(torch.randn(N,C,H,W,generator=g)*0.01).cuda().bfloat16().contiguous(memory_format=torch.channels_last)
    y.backward(dy)
    xf,gf,bf=[t.detach().double().requires_grad_() for t in (x,gm,bt)]
    rf=r.detach().double().requires_grad_() if res else None
    yr=F.batch_norm(xf,None,None,gf,bf,True,0.1,1e-5)
    if res: yr=yr+rf
    if relu: yr=F.relu(yr)
    yr.backward(dy.double())
    rel=lambda a,b:float((a.double()-b).norm()/b.norm())
    print((N,C,H,W,relu,res),"dgamma rel %.2e dbeta rel %.2e dx rel %.2e y rel %.2e"%(rel(gm.grad,gf.grad),rel(bt.grad,bf.grad),rel(x.grad,xf.grad),rel(y,yr)))
