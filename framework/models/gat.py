from gnndelete_b200.models import GAT  # noqa: F401  (reference: framework/models/gat.py)
