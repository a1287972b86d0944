import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mogp_b200 import _capi
from mogp_b200.obsmodel import make_spec
eng=_capi.Engine(0)
spec=make_spec('rbf',True,True,False)
rs=np.random.RandomState(1); m=640
x=np.sort(rs.uniform(0,8,m)); y=(48-6*x+rs.randn(m)*3-30)/20
slot=eng.cluster_create(spec,np.array([0.3,1.0,4.0,0.5]))
eng.cluster_set_data(slot,x,y)
print(eng.cluster_loglik(slot), eng.cluster_debug_stat(slot)[14:16])
