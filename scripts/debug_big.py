import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mogp_b200 import _capi
from mogp_b200.obsmodel import make_spec
eng=_capi.Engine(0)
spec=make_spec('rbf',True,True,False)
for m in [2600]:
    rs=np.random.RandomState(m)
    x=np.sort(rs.uniform(0,8,m)); x[:m//8]=0.0
    y=(48-6*x+rs.randn(m)*3-30)/20
    slot=eng.cluster_create(spec,np.array([0.3,1.0,4.0,0.5]))
    t0=time.time()
    try:
        eng.cluster_set_data(slot,x,y)
        print(m,"ok lml",eng.cluster_loglik(slot),"%.1f ms"%((time.time()-t0)*1e3))
        t0=time.time(); g=eng.cluster_lml_grad(slot); print("   grad",g,"%.1f ms"%((time.time()-t0)*1e3))
    except Exception as e:
        print(m,"FAIL",e)
    eng.cluster_free(slot)
slot=eng.cluster_create(spec,np.array([0.3,1.0,4.0,0.5]))
eng.cluster_set_data(slot,x,y)
st=eng.cluster_debug_stat(slot)
print("k_diag cycles: potrf(+load) %.0f  trtri %.0f" % (st[14], st[15]))
