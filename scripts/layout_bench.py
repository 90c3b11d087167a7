"""Host batch-layout throughput (CPU baseline line 3): the reference's ConvMol.agglomerate_mols, the oracle
restatement and the C++ builder (dcgc_layout_plan/build) on the same zinc-shaped batch of 4096 molecules, one thread each.
The reference is imported from /root/reference with rdkit stubbed (build container only); without it that row is skipped."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np
from deepchem_b200 import mol_graphs as MG
from deepchem_b200.synthetic import make_molecules
from oracle.convmol_layout import OracleConvMol, agglomerate

B = 4096
pm = make_molecules(B, seed=0, shape="zinc")
mols = pm.to_list()


def best(fn, reps):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return min(ts)


rows = []
t = best(lambda: MG.BatchLayout.build(pm), 20)
rows.append({"impl": "C++ builder (dcgc_layout_plan + dcgc_layout_build: reference arrays + CSR / CSR^T / molecule groups / tiles)",
             "ms_per_batch": t * 1e3, "molecules_per_s": B / t})
cms = [OracleConvMol(f, a) for f, a in mols]
t = best(lambda: agglomerate(cms), 3)
rows.append({"impl": "oracle restatement of agglomerate_mols (numpy, ConvMol objects prebuilt)", "ms_per_batch": t * 1e3,
             "molecules_per_s": B / t})
try:
    from _refimport import import_reference
    import_reference()
    from deepchem.feat.mol_graphs import ConvMol
    rcms = [ConvMol(np.asarray(f, dtype=np.float64), a) for f, a in mols]
    t = best(lambda: ConvMol.agglomerate_mols(rcms), 3)
    rows.append({"impl": "reference ConvMol.agglomerate_mols (deepchem/feat/mol_graphs.py:256-349, ConvMol objects prebuilt)",
                 "ms_per_batch": t * 1e3, "molecules_per_s": B / t})
except Exception as e:           # not in the build container
    rows.append({"impl": "reference", "unavailable": str(e)[:100]})
for r in rows:
    r["batch"] = B
    r["cores"] = 1
    print(json.dumps(r), flush=True)
