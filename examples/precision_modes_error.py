import sys, numpy as np, torch
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import nais_testutil as util
from oracle import nais_oracle as orc
from poi_recommendation_models_b200 import ops, synthetic
for style, es in [("reference",None),("trained",None),("trained",1.0)]:
    U,N,beta=6,1500,0.5
    data = synthetic.make_checkins(U,N,seed=21,hist_len=None,max_hist=100,min_hist=5,median_hist=40)
    sd = orc.init_state("region_distance",N,64,64,data.region_num,1,seed=5,style=style)
    if es:
        for k in sd:
            if k.startswith("embed_"): sd[k]=torch.randn(sd[k].shape)*es
    m = util.make_model("region_distance",sd,beta); m.set_catalog(region=data.region,coords=data.coords)
    users = m.make_users(data.indptr,data.indices)
    for prec in ("fp32","tc_split","tc_fast"):
        got = ops.fullrank_scores("region_distance",beta,m._params(),m._catalog,users,precision=prec).cpu().numpy()
        errs=[]
        for u in range(U):
            ref,scale = util.oracle_user_scores(sd,"region_distance",beta,data.coords,data.region,data.history(u),np.arange(N))
            errs.append(util.cond_err(got[u],ref,scale))
        print(style,es,prec,"max cond err %.2e"%max(errs))
