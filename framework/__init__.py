"""Drop-in for the reference's ``framework`` package on the unlearning hot path.

Same import paths and factory signatures as ``/root/reference/framework/__init__.py``
(``get_model`` :51-59, ``get_trainer`` :62-67), backed by the B200 kernels in
``gnndelete_b200``.  Registered: the GNNDelete trainers (edge-logit and node-embedding
objectives, homogeneous and KG) plus the two loops that produce / re-produce the model they
start from (``original``, ``retrain``).  The other baselines
(gradient ascent, Descent-to-Delete, GraphEraser, membership inference …) are out of scope
 and are not registered; asking for one raises ``NotImplementedError``.
"""
from .models import GCN, GAT, GIN, RGCN, GCNDelete, GATDelete, GINDelete, RGCNDelete
from .trainer.base import Trainer, KGTrainer
from .trainer.retrain import RetrainTrainer
from .trainer.gnndelete import GNNDeleteTrainer
from .trainer.gnndelete_nodeemb import GNNDeleteNodeembTrainer, KGGNNDeleteNodeembTrainer

trainer_mapping = {
    'original': Trainer,
    'retrain': RetrainTrainer,
    'gnndelete': GNNDeleteTrainer,
    'gnndelete_mse': GNNDeleteTrainer,
    'gnndelete_nodeemb': GNNDeleteNodeembTrainer,
}

kg_trainer_mapping = {
    'original': KGTrainer,            # eval / test only: KG training of the original model is not accelerated
    'gnndelete': KGGNNDeleteNodeembTrainer,
    'gnndelete_nodeemb': KGGNNDeleteNodeembTrainer,
}


def get_model(args, mask_1hop=None, mask_2hop=None, num_nodes=None, num_edge_type=None):
    if 'gnndelete' in args.unlearning_model:
        mapping = {'gcn': GCNDelete, 'gat': GATDelete, 'gin': GINDelete, 'rgcn': RGCNDelete}
    else:
        mapping = {'gcn': GCN, 'gat': GAT, 'gin': GIN, 'rgcn': RGCN}
    if args.gnn not in mapping:
        raise NotImplementedError(f'gnn={args.gnn!r} is outside the accelerated hot path')
    return mapping[args.gnn](args, mask_1hop=mask_1hop, mask_2hop=mask_2hop, num_nodes=num_nodes,
                             num_edge_type=num_edge_type)


def get_trainer(args):
    mapping = kg_trainer_mapping if args.gnn in ['rgcn', 'rgat'] else trainer_mapping
    if args.unlearning_model not in mapping:
        raise NotImplementedError(f'unlearning_model={args.unlearning_model!r} is outside the accelerated hot path')
    return mapping[args.unlearning_model](args)
