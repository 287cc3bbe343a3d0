"""TEST INFRASTRUCTURE — CPU restatement of the reference's unlearning driver
arithmetic: the mask-building preamble of ``delete_gnn.py`` and the loss bodies
of ``framework/trainer/gnndelete.py`` / ``gnndelete_nodeemb.py``.
PARITY UNPINNED (see ``oracle/pyg_ops.py``).  Never imported by the product.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from gnndelete_b200.data import GraphData

from . import pyg_ops as P


# ------------------------------------------------------------------ mask pipeline
def build_unlearning_data(data, df_mask, num_edge_type=None):
    """``delete_gnn.py:113-189`` given the sampled directed ``df_mask``.

    Homogeneous: k-hop masks on the DIRECTED list, then ``to_undirected`` carrying
    the two masks.  KG (``num_edge_type`` set): reverse edges are concatenated with
    types ``+num_edge_type`` and the masks are ``repeat(2)``-ed, no sort
    (:158-172)."""
    out = data.clone()
    ei = data.train_pos_edge_index
    n = data.num_nodes
    out.directed_df_edge_index = ei[:, df_mask]                            # :113
    if num_edge_type is not None:
        out.directed_df_edge_type = data.train_edge_type[df_mask]          # :115
    seeds = ei[:, df_mask].flatten().unique()
    _, two_hop_edge, _, two_hop_mask = P.k_hop_subgraph(seeds, 2, ei, num_nodes=n)   # :128-132
    _, one_hop_edge, _, _ = P.k_hop_subgraph(seeds, 1, ei, num_nodes=n)             # :136-140
    s1 = torch.zeros(n, dtype=torch.bool)
    s2 = torch.zeros(n, dtype=torch.bool)
    s1[one_hop_edge.flatten().unique()] = True                              # :144
    s2[two_hop_edge.flatten().unique()] = True                              # :145
    out.sdf_node_1hop_mask = s1
    out.sdf_node_2hop_mask = s2
    dr_mask = ~df_mask
    if num_edge_type is not None:                                           # :158-172
        r, c = ei
        out.edge_index = torch.cat([ei, torch.stack([c, r], 0)], 1)
        out.edge_type = torch.cat([data.train_edge_type, data.train_edge_type + num_edge_type], 0)
        two_hop_mask = two_hop_mask.repeat(2).view(-1)
        df_mask = df_mask.repeat(2).view(-1)
        dr_mask = dr_mask.repeat(2).view(-1)
    else:                                                                   # :175-182
        sym, (df_i, sdf_i) = P.to_undirected(ei, [df_mask.int(), two_hop_mask.int()])
        two_hop_mask = sdf_i.bool()
        df_mask = df_i.bool()
        dr_mask = ~df_mask
        out.train_pos_edge_index = sym
        out.edge_index = sym
    out.sdf_mask = two_hop_mask                                             # :187-189
    out.df_mask = df_mask
    out.dr_mask = dr_mask
    return out


def dense_pair_mask(data, node_mask):
    """``gnndelete.py:163-193``: all node pairs inside the S_Df node set, minus the
    Df pairs, strictly-lower-triangular.  O(N^2) — small graphs only."""
    n = data.num_nodes
    m = node_mask.view(-1, 1) & node_mask.view(1, -1)
    df = data.train_pos_edge_index[:, data.df_mask]
    m[df[0], df[1]] = False
    m[df[1], df[0]] = False
    return m & torch.ones(n, n, dtype=torch.bool).tril(-1)


# ----------------------------------------------------------------- epoch bodies
def fullbatch_loss(model, data, neg_edge_index, logits_ori, pair_mask):
    """``train_fullbatch`` epoch body, ``gnndelete.py:215-250`` with supplied
    negatives.  Returns (loss, loss_r, loss_l, z)."""
    z = model(data.x, data.train_pos_edge_index[:, data.sdf_mask])
    n = int(data.df_mask.sum())
    df_logits = model.decode(z, data.train_pos_edge_index[:, data.df_mask], neg_edge_index)
    loss_r = F.mse_loss(df_logits[:n], df_logits[n:])
    if int(pair_mask.sum()) != 0:
        loss_l = F.mse_loss((z @ z.t())[pair_mask].sigmoid(), logits_ori[pair_mask].sigmoid())
    else:
        loss_l = torch.zeros(())
    return 0.5 * loss_r + 0.5 * loss_l, loss_r, loss_l, z


def edge_form_loss(model, data, neg_edge_index, z_ori, masks_positional=True):
    """``train_minibatch`` step body on the whole graph, ``gnndelete.py:352-398``:
    Randomness MSE on (Df, neg) logits + edge-form Neighbourhood-Influence MSE over
    the S_Df edges with ``u < v`` (no sigmoid).  Returns (loss, loss_r, loss_l, z)."""
    ei = data.train_pos_edge_index
    if masks_positional:                                                    # :352
        z = model(data.x, ei[:, data.sdf_mask], data.sdf_node_1hop_mask, data.sdf_node_2hop_mask)
    else:
        z = model(data.x, ei[:, data.sdf_mask])
    n = int(data.df_mask.sum())
    df_logits = model.decode(z, ei[:, data.df_mask], neg_edge_index)       # :362
    loss_r = F.mse_loss(df_logits[:n], df_logits[n:])                       # :363
    edge = ei[:, data.sdf_mask]                                             # :379
    lower = edge[0] < edge[1]
    row, col = edge[0][lower], edge[1][lower]
    logits_ori = (z_ori[row] * z_ori[col]).sum(-1)                          # :383
    logits = (z[row] * z[col]).sum(-1)
    loss_l = F.mse_loss(logits, logits_ori)                                 # :386
    return 0.5 * loss_r + 0.5 * loss_l, loss_r, loss_l, z


def negative_sampling_kg(edge_index, edge_type, generator=None):
    """``framework/utils.py:46-58``: per relation, permute the heads."""
    out = edge_index.clone()
    for et in edge_type.unique():
        m = edge_type == et
        old = out[0, m]
        out[0, m] = old[torch.randperm(old.shape[0], generator=generator)]
    return out


def kg_non_df_masks(data):
    """``gnndelete_nodeemb.py:723-727``."""
    non_df = torch.ones(data.x.shape[0], dtype=torch.bool)
    non_df[data.directed_df_edge_index.flatten().unique()] = False
    return data.sdf_node_1hop_mask & non_df, data.sdf_node_2hop_mask & non_df


def kg_step_losses(model, data, neg_edge_index, num_edge_type, alpha=0.5):
    """``KGGNNDeleteNodeembTrainer.train`` step body on the whole graph,
    ``gnndelete_nodeemb.py:749-798``.  Returns (loss1, loss2, parts) — the caller
    performs the two backward / Adam steps (:788-796)."""
    m1, m2 = kg_non_df_masks(data)
    edge_index = data.edge_index[:, data.dr_mask]                           # :749-750
    edge_type = data.edge_type[data.dr_mask]
    z1, z2 = model(data.x, edge_index, edge_type, m1, m2, return_all_emb=True)
    with torch.no_grad():                                                   # :754-755
        z1o, z2o = model.get_original_embeddings(data.x, edge_index, edge_type, return_all_emb=True)
    pos_ei = data.edge_index[:, data.df_mask]                               # :758-763
    pos_et = data.edge_type[data.df_mask]
    dec = pos_et < num_edge_type
    dec_ei = pos_ei[:, dec]
    e1 = torch.cat([z1[dec_ei[0]], z1[dec_ei[1]]], 0)                        # :770-774
    e1o = torch.cat([z1o[neg_edge_index[0]], z1o[neg_edge_index[1]]], 0)
    e2 = torch.cat([z2[dec_ei[0]], z2[dec_ei[1]]], 0)
    e2o = torch.cat([z2o[neg_edge_index[0]], z2o[neg_edge_index[1]]], 0)
    loss_r1 = F.mse_loss(e1, e1o)                                           # :776-777
    loss_r2 = F.mse_loss(e2, e2o)
    loss_l1 = F.mse_loss(z1[m1], z1o[m1])                                   # :780-781
    loss_l2 = F.mse_loss(z2[m2], z2o[m2])
    loss1 = alpha * loss_r1 + (1 - alpha) * loss_l1                         # :788
    loss2 = alpha * loss_r2 + (1 - alpha) * loss_l2                         # :793
    return loss1, loss2, dict(loss_r1=loss_r1, loss_r2=loss_r2, loss_l1=loss_l1, loss_l2=loss_l2,
                              z1=z1, z2=z2, decoding_edge_index=dec_ei, decoding_edge_type=pos_et[dec])
