import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bioshell_b200 as bs
from bioshell_b200 import clustering as cl
from oracle import c_oracle
rng = np.random.default_rng(0)
a = np.tril(rng.integers(1, 9, (70, 70)).astype(np.float32), -1)
m = a + a.T
with bs.Context(0) as ctx:
    for link, name in ((cl.average_link, "average"), (cl.single_link, "single"), (cl.median_link, "median")):
        ref = c_oracle.hclust(m, name)
        for rep in range(2):
            mi, mj, md = cl.hclust_merge_log(70, m, link, ctx)
            bi = np.nonzero((mi != ref["mat_i"]) | (mj != ref["mat_j"]))[0]
            bd = np.nonzero(md != ref["dist"])[0]
            print(name, "rep", rep, "index mismatches", len(bi), "dist mismatches", len(bd),
                  [(int(s), float(md[s]), float(ref["dist"][s])) for s in bd[:6]])
