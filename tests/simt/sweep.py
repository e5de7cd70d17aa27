"""TEST INFRASTRUCTURE ONLY — every regret x weight x sampling combination of the flat-game training kernels (Kuhn: all 80 at two batch shapes, Leduc: a third)
under the SIMT shim against the oracle; 212 cases, ~2.5 min.  `python tests/simt/sweep.py` (tests/test_simt_mccfr.py keeps a short subset in the suite)."""
import sys, ctypes, time, itertools
import os; ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from oracle import binding as oracle
from simt import build as sb
ROW = oracle.ROW_DTYPE
l = ctypes.CDLL(sb.build())
vp,u64,i32 = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int
l.simt_create.restype = vp; l.simt_create.argtypes=[i32]*5+[u64]
l.simt_destroy.argtypes=[vp]; l.simt_step.argtypes=[vp,u64]; l.simt_export.argtypes=[vp,i32,vp,i32]
S={"ExternalSampling":0,"PrunableSampling":2,"PluribusSampling":3,"TargetedSampling":4}
def rows(h):
    buf=np.zeros(4096,dtype=ROW); n=l.simt_export(h, 0, buf.ctypes.data, len(buf)); return buf[:n]
bad=0; n=0; t0=time.time()
for game,batch,steps in (("kuhn",1,120),("kuhn",130,5),("leduc",1,10),("leduc",300,3)):
    for reg in oracle.REGRETS:
        for wt in oracle.WEIGHTS:
            for smp in S:
                if game=="leduc" and (hash((reg,wt,smp))%3): continue   # a third of the combinations on Leduc
                h=l.simt_create(oracle.GAMES[game], oracle.REGRETS[reg], oracle.WEIGHTS[wt], S[smp], batch, 5)
                o=oracle.OracleSolver(game,reg,wt,smp,batch=batch,seed=5)
                l.simt_step(h,steps); o.step(steps)
                ok = rows(h).tobytes()==o.profile_rows().tobytes()
                n+=1
                if not ok: bad+=1; print("MISMATCH",game,reg,wt,smp,batch)
                l.simt_destroy(h)
print("cases",n,"bad",bad,"%.0fs"%(time.time()-t0))
