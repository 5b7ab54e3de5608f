"""Debug helper: first field / cell / step where the CUDA SVAT kernel and the CPU oracle (roger_pow mode) differ in any bit."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle.engine import OracleSvatEngine, lib
from roger_b200 import synth, abi
from roger_b200.engine import CudaSvatEngine
from roger_b200.svat import SvatModel
lib().roger_oracle_set_pow(1)
lateral = (sys.argv[1] == "oned") if len(sys.argv) > 1 else True
doy = int(sys.argv[2]) if len(sys.argv) > 2 else 121
nsteps = int(sys.argv[3]) if len(sys.argv) > 3 else 60
TILE=(32,32)
forcing = synth.mixed_forcing(40, seed=42, start_doy=doy)
prim, th_rz, th_ss = synth.heterogeneous_parameters(TILE, seed=43, lateral=lateral, lu_choices=(10, 11, 12, 8, 5, 9, 6))
tile = SvatModel(CudaSvatEngine, TILE, forcing, prim, th_rz, th_ss, lateral=lateral)
ref = SvatModel(OracleSvatEngine, TILE, forcing, prim, th_rz, th_ss, lateral=lateral)
for m in (tile, ref): m.setup()
tile.engine.set_block_stepping(False)
kinds=dict(zip(abi.FIELD_NAMES, abi.FIELD_KINDS))
for step in range(nsteps):
    a=tile.step(True); b=ref.step(True)
    bad=[]
    for n in abi.FIELD_NAMES:
        x=tile.get(n); y=ref.get(n)
        ne=~((x==y)|(np.isnan(x)&np.isnan(y)))
        if ne.any():
            i=int(np.argmax(ne)); bad.append((n,int(ne.sum()),i,float(x[i]),float(y[i])))
    if bad:
        print('step',step+1,'dt',a.dt_secs,'fields differing:',len(bad)); 
        for t in bad[:40]: print('  ',t)
        break
else:
    print('bit-identical over',nsteps,'steps,',TILE[0]*TILE[1],'cells, lateral',lateral,'doy',doy)
