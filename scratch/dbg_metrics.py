import sys, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/oracle'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
import oracle_lib
from besst_b200 import abi, libmetrics
from besst_b200.records import RecordBatch
from besst_b200.engine import CudaEngine
batch = RecordBatch.load('/root/repo/tests/golden/testset1_head.npz')
params = abi.make_params("fr", 11, 100.0, 0.0, 0.0, 0.0)
rows = libmetrics.metric_rows(batch.lengths)
eng = CudaEngine()
for name, f in (("oracle", oracle_lib.libmetrics), ("cuda", eng.libmetrics)):
    rc, m, adj = f(rows, params, batch, batch.lengths, True)
    print(name, rc, {k: getattr(m, k) for k, _ in m._fields_}, adj.sum(), len(adj))
