from gnndelete_b200.trainer import GNNDeleteTrainer, get_loss_fct  # noqa: F401  (reference: framework/trainer/gnndelete.py)
