"""CTC gradient of the kernel and of torch fp32 against an fp64 reference at growing sizes (precision budget of the fp32 recursions)."""
import sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import torch, numpy as np, torch.nn.functional as F
from oracle import model_ref as M, decode_np as D
from nb_asr_b200 import _lib
import gpu_utils as U
lib=_lib.load()
for T,S in [(64,20),(300,100),(700,250),(700,270),(700,300)]:
    torch.manual_seed(11)
    B,V=4,49
    logits=(torch.randn(B,T,V)*2).double().requires_grad_(True)
    logp=F.log_softmax(logits,2)
    alen=torch.tensor([4*T,4*(T-3),4*(T//2)+1,4*max(2,S//4)])
    tl=torch.tensor([S,S-3,0,S])
    tg=torch.randint(1,V,(B,S),dtype=torch.int32)
    tg[0,1::2]=tg[0,0:-1:2][:tg[0,1::2].numel()]
    for b in range(B): tg[b,int(tl[b]):]=0
    out_len=alen//4
    loss=M.ctc_loss_ref(logp,out_len,tg,tl)            # fp64 reference
    gl,=torch.autograd.grad(loss,logits)
    loss32=M.ctc_loss_ref(logp.float(),out_len,tg,tl)
    logits32=logits.detach().float().requires_grad_(True)
    l32=M.ctc_loss_ref(F.log_softmax(logits32,2),out_len,tg,tl); g32,=torch.autograd.grad(l32,logits32)
    lp=logp.detach().float().to(U.DEV).contiguous()
    tgd,alend,tld=tg.to(U.DEV),alen.to(U.DEV),tl.to(U.DEV)
    nll=torch.zeros(B,device=U.DEV); lossd=torch.zeros(1,device=U.DEV); dl=torch.zeros(B,T,V,device=U.DEV)
    work=torch.zeros(2*B*T*(2*S+1)+16,device=U.DEV)
    _lib.check(lib.nbasr_ctc(lp.data_ptr(),B,T,V,tgd.data_ptr(),S,alend.data_ptr(),4,tld.data_ptr(),nll.data_ptr(),lossd.data_ptr(),dl.data_ptr(),work.data_ptr(),U.stream()))
    torch.cuda.synchronize()
    per=[float((dl[b].cpu().double()-gl[b]).norm()/gl[b].norm().clamp_min(1e-30)) for b in range(3)]
    print(T,S,'kernel vs fp64:',['%.1e'%x for x in per],' torch fp32 vs fp64: %.1e'%float((g32.double()-gl).norm()/gl.norm()), ' loss rel %.1e'%abs(lossd.item()-float(loss))/abs(float(loss)) if False else '')
