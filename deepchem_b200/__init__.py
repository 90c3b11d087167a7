"""B200-native GraphConv / GraphPool / GraphGather / DMPNN hot path behind DeepChem's API."""
__version__ = "0.1.0"
