"""CPU emulation (numpy float16 rounding, float32 accumulation) of the tensor path's arithmetic: single-pass fp16 (x1), single-pass
main rows with split S/L rows (x1sl = NAIS_PREC_TC_FAST) and the three-pass hi/lo split (x3 = NAIS_PREC_TC_SPLIT), against the
float64 oracle, at several weight scales.  This is the study behind the precision scheme in DESIGN.md section 3."""
import numpy as np, torch, sys
sys.path.insert(0,'.')
from oracle import nais_oracle as orc
from poi_recommendation_models_b200 import synthetic
def split16(x):
    hi = x.astype(np.float16); lo = (x - hi.astype(np.float32)).astype(np.float16); return hi, lo
def p2floor(x): return 2.0**np.floor(np.log2(x))
def run(style, N=2000, D=64, hid=64, H=64, B=256, seed=0, emb_std=None, wscale=1.0):
    rng=np.random.default_rng(seed)
    coords, region, R = synthetic.make_catalog(N, seed=1)
    sd = orc.init_state("region_distance", N, D, hid, R, 1, seed=3, style=style)
    if emb_std is not None:
        for k in sd:
            if k.startswith('embed_'): sd[k] = torch.randn(sd[k].shape)*emb_std
    sd['attn_layer1.weight'] *= wscale; sd['attn_layer2.weight'] *= wscale
    hist = np.stack([rng.choice(N,H,replace=False) for _ in range(B)]); tgt = rng.integers(0,N,B)
    aux = orc.latlon_abs_diff(coords,tgt,hist)
    t=lambda a: torch.from_numpy(a)
    ref, scale = orc.attention_network_with_scale(sd,"region_distance",0.5,t(hist),t(tgt),t(region[hist]),t(region[tgt]),t(aux),dtype=torch.float64)
    ref32 = orc.attention_network(sd,"region_distance",0.5,t(hist),t(tgt),t(region[hist]),t(region[tgt]),t(aux),dtype=torch.float32).numpy()
    ref=ref.numpy(); scale=scale.numpy()
    P={k:v.numpy().astype(np.float32) for k,v in sd.items()}
    q = np.concatenate([P['embed_history.weight'][hist], P['embed_region.weight'][region[hist]]],-1)  # B,H,D
    p = np.concatenate([P['embed_target.weight'][tgt], P['embed_region.weight'][region[tgt]]],-1)   # B,D
    W=P['attn_layer1.weight']; b=P['attn_layer1.bias']; v=P['attn_layer2.weight'][0]
    Wd=P['dist_layer.weight']; bd=P['dist_layer.bias']
    z = (aux*np.float32(100.0))@Wd.T + bd; g = (1/(1+np.exp(-z))).astype(np.float32)  # B,H,2
    c = 0.5*np.abs(v); sign=np.sign(v)
    maxP=np.abs(p).max(); maxQ=np.abs(q).max(); CW = c[:,None]*W[:,:D]; u=(sign[:,None]*CW).sum(0)
    maxB = maxQ*max(np.abs(CW).max(), np.abs(u).max())
    sA=p2floor(512/maxP); sS=p2floor(512/maxQ); sB=p2floor(512/maxB); sig=sA*sB; sAe=256.0; sBe=sig/sAe
    ext_main = np.concatenate([c[:,None]*W[:,D:], (c*b)[:,None]],1)  # hid,3
    ext_L = (sign[:,None]*ext_main).sum(0)
    maxBe=max(np.abs(ext_main).max(), np.abs(ext_L).max())
    while maxBe*sBe > 2**14: sB/=2; sig=sA*sB; sBe=sig/sAe
    out={}
    for mode in ('x1','x1sl','x3'):
        # A: [B, D+3] per (b,h): p*sA, g*sAe, 1*sAe ; Bop: per (b,h): rows hid+2, cols D+3
        A_x = (p*sA).astype(np.float32); A_e = np.concatenate([g*sAe, np.full(g.shape[:2]+(1,),sAe,np.float32)],-1)
        Bx_main = (CW[None,None]*q[:,:,None,:]*sB).astype(np.float32)   # B,H,hid,D
        Bx_L = (u[None,None]*q*sB).astype(np.float32)                    # B,H,D
        Bx_S = (q*sS).astype(np.float32)
        Be_main=(ext_main*sBe).astype(np.float32); Be_L=(ext_L*sBe).astype(np.float32)
        def mm(Ah,Al,Bh,Bl,eq,aux=False):
            f=lambda a:a.astype(np.float32)
            r = np.einsum(eq,f(Ah),f(Bh))
            if mode=='x3' or (mode=='x1sl' and aux): r = r + np.einsum(eq,f(Ah),f(Bl)) + np.einsum(eq,f(Al),f(Bh))
            return r.astype(np.float32)
        Axh,Axl=split16(A_x); Aeh,Ael=split16(A_e)
        Bmh,Bml=split16(Bx_main); BLh,BLl=split16(Bx_L); BSh,BSl=split16(Bx_S); Bemh,Beml=split16(Be_main); BeLh,BeLl=split16(Be_L)
        tmain = mm(Axh,Axl,Bmh,Bml,'bd,bhkd->bhk') + mm(Aeh,Ael,Bemh,Beml,'bhe,ke->bhk')
        L = mm(Axh,Axl,BLh,BLl,'bd,bhd->bh',True) + mm(Aeh,Ael,BeLh,BeLl,'bhe,e->bh',True)
        S = mm(Axh,Axl,BSh,BSl,'bd,bhd->bh',True)
        acc = (np.abs(tmain)*sign[None,None]).sum(-1,dtype=np.float32)
        a = ((L+acc)*np.float32(1/sig)).astype(np.float32); s=(S*np.float32(1/(sA*sS))).astype(np.float32)
        e = np.exp(a)*(hist!=tgt[:,None]); 
        score = (e*s).sum(-1,dtype=np.float32)/np.sqrt(e.sum(-1,dtype=np.float32))
        out[mode]=np.max(np.abs(score-ref)/np.maximum(np.abs(ref),scale))
    out['ref32']=np.max(np.abs(ref32-ref)/np.maximum(np.abs(ref),scale))
    a64 = None
    return out
for style,es,ws in [('reference',None,1),('trained',None,1),('trained',1.0,1),('trained',0.3,4.0),('trained',1.0,4.0)]:
    print(style,es,ws, run(style,emb_std=es,wscale=ws))
