from . import deletion, gat, gcn, gin, rgcn  # noqa: F401  (reference module paths)
from gnndelete_b200.models import (GCN, GAT, GIN, RGCN, GCNDelete, GATDelete, GINDelete, RGCNDelete,  # noqa: F401
                                   DeletionLayer)
