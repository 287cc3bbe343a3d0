from gnndelete_b200.trainer import KGGNNDeleteNodeembTrainer  # noqa: F401  (reference: framework/trainer/gnndelete_nodeemb.py)
