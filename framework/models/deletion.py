from gnndelete_b200.models import (DeletionLayer, GCNDelete, GATDelete, GINDelete,  # noqa: F401
                                   RGCNDelete)
