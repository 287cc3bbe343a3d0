from gnndelete_b200.models import (GCN, GAT, GIN, RGCN, GCNDelete, GATDelete, GINDelete, RGCNDelete,
                                   DeletionLayer)
