from gnndelete_b200.trainer import Trainer, KGTrainer  # noqa: F401  (reference: framework/trainer/base.py)
