from gnndelete_b200.models import RGCN  # noqa: F401  (reference: framework/models/rgcn.py)
