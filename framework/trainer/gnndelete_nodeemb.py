# reference: framework/trainer/gnndelete_nodeemb.py
from gnndelete_b200.trainer import GNNDeleteNodeembTrainer, KGGNNDeleteNodeembTrainer  # noqa: F401
from gnndelete_b200.trainer import get_nodeemb_loss_fct as get_loss_fct  # noqa: F401
