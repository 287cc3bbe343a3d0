from gnndelete_b200.models import GCN  # noqa: F401  (reference: framework/models/gcn.py)
