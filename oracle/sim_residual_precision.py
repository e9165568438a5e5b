"""TEST INFRASTRUCTURE (part of the oracle): CPU simulation of the bf16 encoder's roundings to decide how the residual stream may be
stored.  Runs the whole encoder of one model on one synthetic clip with bf16 GEMM operands, the LayerNorm folded into the consuming
GEMM (as csrc/gemm_tc.cu does) and the residual stream rounded to a chosen dtype after every update, then the fp32 TL-TR head, and
prints the max |logit - fp32 reference| and the pooled states' relative error.

    python oracle/sim_residual_precision.py large-v2        # ~1 min per variant on 8 cores
Result quoted in DESIGN.md §3 (large-v2): fp32 residual 5.5e-3, fp16 5.4e-3, bf16 5.8e-3 (pooled 3.1e-3 / 3.3e-3 / 1.1e-2)."""
import os, sys, math, time, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0]=[os.path.join(ROOT, 'whisper-at_b200'), os.path.join(ROOT, 'oracle')]
import wat_oracle as O
from whisper_at import synth
import torch.nn.functional as F
torch.set_num_threads(8)
name=sys.argv[1] if len(sys.argv)>1 else 'small'
n_mels=128 if name=='large-v2' else 80
low = name in ('small','medium')
d,h,L=synth.MODEL_SHAPES[name]
sd=synth.synth_state_dict(n_mels,d,L,low,seed=1,init='lively')
clip=synth.synth_clip(1)
mel=O.log_mel_clip(clip,n_mels)
bf=lambda t: t.bfloat16().float()
def block(x, p, n_head, res_dtype, a_dtype):
    g=lambda k: sd[f"{p}.{k}"]
    rnd=lambda t: t.to(res_dtype).float()
    N,T,D=x.shape
    def ln_gemm(x, lnp, wk, bk):
        # folded LN: A operand = x rounded to a_dtype (un-normalised), stats from fp32 x (before residual rounding) ~ use x itself
        xa = x.to(a_dtype).float()
        mean = x.mean(-1,keepdim=True); var = x.var(-1,unbiased=False,keepdim=True); rstd=(var+1e-5).rsqrt()
        Wp = bf(g(wk+'.weight')*g(lnp+'.weight'))
        cs = Wp.sum(-1)
        bp = (g(bk+'.bias') if (bk+'.bias') in [k[len(p)+1:] for k in sd if k.startswith(p)] else 0) + g(wk+'.weight')@g(lnp+'.bias')
        return rstd*(xa@Wp.T - mean*cs) + bp
    q=ln_gemm(x,'attn_ln','attn.query','attn.query'); k=ln_gemm(x,'attn_ln','attn.key','attn.keyNOBIAS'); v=ln_gemm(x,'attn_ln','attn.value','attn.value')
    hd=D//n_head
    q=bf(q*0.125*1.4426950408889634); k=bf(k); v=bf(v)
    qh=q.view(N,T,n_head,hd).permute(0,2,1,3); kh=k.view(N,T,n_head,hd).permute(0,2,3,1); vh=v.view(N,T,n_head,hd).permute(0,2,1,3)
    s=qh@kh
    P=torch.exp2(s - s.amax(-1,keepdim=True)); l=P.sum(-1,keepdim=True)
    a=bf((bf(P)@vh)/l).permute(0,2,1,3).reshape(N,T,D)
    x = rnd(x + a@bf(g('attn.out.weight')).T + g('attn.out.bias'))
    hmid = bf(O._gelu(ln_gemm(x,'mlp_ln','mlp.0','mlp.0')))
    return rnd(x + hmid@bf(g('mlp.2.weight')).T + g('mlp.2.bias'))
def run(res_dtype, a_dtype):
    x=bf(mel)[None]
    x=bf(O._gelu(F.conv1d(x, bf(sd['encoder.conv1.weight']), sd['encoder.conv1.bias'], padding=1)))
    x=O._gelu(F.conv1d(x, bf(sd['encoder.conv2.weight']), sd['encoder.conv2.bias'], stride=2, padding=1)).permute(0,2,1)+O.sinusoids(1500,d)
    x=x.to(res_dtype).float()
    pooled=[]
    for i in range(L):
        x=block(x,f'encoder.blocks.{i}',h,res_dtype,a_dtype)
        pooled.append(x.to(a_dtype).float().reshape(1,75,20,d).mean(2))
    return torch.stack(pooled,1)
with torch.no_grad():
    t0=time.time()
    ref=O.encoder_pooled(mel[None],sd,h)
    lg_ref=O.tltr_head(ref,sd,10)
    for res_dtype,a_dtype,tag in ((torch.float32,torch.bfloat16,'fp32 residual, bf16 A operand (round 1)'),(torch.float16,torch.bfloat16,'fp16 residual, bf16 A operand (round 2)'),(torch.bfloat16,torch.bfloat16,'bf16 residual = A operand')):
        p=run(res_dtype,a_dtype)
        lg=O.tltr_head(p,sd,10)
        print(name,tag,'pooled relerr',float((p-ref).abs().max()/ref.abs().max()),'logit maxabs',float((lg-lg_ref).abs().max()), 'top5 same', [set(torch.topk(a,5).indices.tolist())==set(torch.topk(b,5).indices.tolist()) for a,b in zip(lg[0],lg_ref[0])], f'{time.time()-t0:.0f}s',flush=True)
