import sys, os
sys.path.insert(0, '/root/repo')
import torch, torch.nn.functional as F
from pixelssl_b200 import ops
from oracle import adv_oracle as A
CL = torch.channels_last
st = A.init_fcd(5)
g = torch.Generator().manual_seed(3)
prob = torch.softmax(torch.randn(2, 21, 65, 65, generator=g), 1)
def rel(a,b):
    a,b=a.detach().double().cpu(),b.detach().double().cpu(); return float((a-b).abs().max()/b.abs().max().clamp_min(1e-30))
# CPU chain with retained grads
xs=[prob.clone().requires_grad_(True)]
x=xs[0]
acts=[]
for name in A.FCD_LAYERS[:-1]:
    c=F.conv2d(x, st[name+'.weight'], st[name+'.bias'], stride=2, padding=1); c.retain_grad()
    x=F.leaky_relu(c,0.2); x.retain_grad(); acts.append((c,x))
c=F.conv2d(x, st['classifier.weight'], st['classifier.bias'], stride=2, padding=1); c.retain_grad()
out=F.interpolate(c,size=(65,65),mode='bilinear',align_corners=True)
w=torch.randn(out.shape, generator=g)
(out*w).sum().backward()
# GPU chain
from pixelssl_b200.nn.modules import Conv2d
pg=prob.cuda().requires_grad_(True)
xg=ops.planar_to_nhwc(pg); xg.retain_grad()
gacts=[]
x=xg
convs=[]
chans=[21,64,128,256,512,1]
for i,name in enumerate(A.FCD_LAYERS):
    m=Conv2d(chans[i],chans[i+1],4,stride=2,padding=1).cuda()
    m.weight.data.copy_(st[name+'.weight']); m.bias.data.copy_(st[name+'.bias']); convs.append(m)
for m in convs[:-1]:
    cg=m(x); cg.retain_grad(); x=ops.leaky_relu(cg,0.2); x.retain_grad(); gacts.append((cg,x))
cg=convs[-1](x); cg.retain_grad()
og=ops.bilinear(cg,(65,65),True,channels=1,nhwc=True)
(og*w.cuda()).sum().backward()
print('out', rel(og,out), 'dcls', rel(cg.grad, c.grad))
for i in range(3,-1,-1):
    print('layer',i,'fwd conv',rel(gacts[i][0],acts[i][0]),'d_act',rel(gacts[i][1].grad,acts[i][1].grad),'d_conv',rel(gacts[i][0].grad,acts[i][0].grad))
print('d_in_nhwc', rel(xg.grad[:, :21], xs[0].grad), 'pg', rel(pg.grad, xs[0].grad))
d=(pg.grad.cpu()-xs[0].grad).abs(); idx=d.flatten().argmax(); print('worst idx', torch.unravel_index(idx, d.shape), float(d.max()), float(xs[0].grad.abs().max()))
for i in range(4):
    cg_, c_ = gacts[i][0].detach().cpu(), acts[i][0].detach()
    flips = ((cg_ > 0) != (c_ > 0))
    print('layer', i, 'sign flips', int(flips.sum()), 'min|c| cpu', float(c_.abs().min()), 'where flips |c|:', c_[flips].abs().tolist()[:5])
