from gnndelete_b200.models import GIN  # noqa: F401  (reference: framework/models/gin.py)
