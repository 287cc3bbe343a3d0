from gnndelete_b200.trainer import RetrainTrainer  # noqa: F401  (reference: framework/trainer/retrain.py)
