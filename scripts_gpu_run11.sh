#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 500 -k dense_tiles 2>&1 | grep -E "assert|Error|shape|^E" | head -20
python - <<'PY'
import numpy as np, scipy.sparse as sp, sys
sys.path.insert(0,'.')
import single_algebra_b200 as s
ctx=s.default_context()
rng=np.random.default_rng(9)
for shape,dens in (((300,200),1.0),((257,130),0.6),((5,3),1.0),((1,70),0.5),((130,64),1.0),((128,64),0.2)):
    D=rng.integers(1,9,size=shape).astype(np.float32)*(rng.random(shape)<dens); D[0,0]=3
    A=sp.csr_matrix(D); d=s.CsrMatrix.from_scipy(A,ctx).to_device()
    for tr in (False,True):
        X=rng.standard_normal((shape[0] if tr else shape[1],60)).astype(np.float32)
        ref=(A.T if tr else A).astype(np.float64)@X.astype(np.float64)
        got=s.op_spmm(d,X,transposed=tr)
        err=np.abs(got-ref).max()/max(np.abs(ref).max(),1e-30)
        bad=np.argwhere(np.abs(got-ref)>1e-3*np.abs(ref).max())
        print(shape,dens,tr,'err',err,'n_bad',len(bad), bad[:3].tolist())
PY
