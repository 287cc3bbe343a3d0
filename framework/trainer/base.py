from gnndelete_b200.trainer import Trainer  # noqa: F401  (reference: framework/trainer/base.py)
