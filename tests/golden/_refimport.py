"""Import the reference deepchem from /root/reference without RDKit / TF / PyG.

Only used by the golden-vector generator scripts in this directory (run in the build
container, where /root/reference exists).  Nothing in tests/, bench.py or the product
imports this at run time on the GPU box.
"""
import sys
import types
from unittest.mock import MagicMock

_RDKIT = [
    "rdkit", "rdkit.Chem", "rdkit.Chem.AllChem", "rdkit.Chem.rdchem", "rdkit.Chem.Draw",
    "rdkit.Chem.rdmolops", "rdkit.Chem.rdMolTransforms", "rdkit.Chem.Descriptors",
    "rdkit.DataStructs", "rdkit.Chem.rdmolfiles", "rdkit.Chem.rdChemReactions",
    "rdkit.Chem.rdFingerprintGenerator", "rdkit.RDLogger", "rdkit.Chem.rdPartialCharges",
    "rdkit.Chem.Scaffolds", "rdkit.Chem.Scaffolds.MurckoScaffold", "rdkit.ML",
    "rdkit.ML.Cluster", "rdkit.Chem.rdMolDescriptors", "rdkit.Geometry",
    "rdkit.Chem.rdDistGeom", "rdkit.Chem.rdForceFieldHelpers",
]


def import_reference(path="/root/reference"):
    for name in _RDKIT:
        sys.modules.setdefault(name, MagicMock())
    if "torch_geometric" not in sys.modules:
        tg = types.ModuleType("torch_geometric")
        tgd = types.ModuleType("torch_geometric.data")

        class Data(object):
            def __init__(self, **kw):
                for k, v in kw.items():
                    setattr(self, k, v)

            def __inc__(self, key, value, *a, **k):
                return 0

        class Batch(object):
            pass

        tgd.Data = Data
        tgd.Batch = Batch
        tg.data = tgd
        sys.modules["torch_geometric"] = tg
        sys.modules["torch_geometric.data"] = tgd
    if path not in sys.path:
        sys.path.insert(0, path)
    sys.dont_write_bytecode = True
    import deepchem  # noqa: F401
    return deepchem
