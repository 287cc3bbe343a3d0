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


def _centering(K):
    """``gnndelete_nodeemb.py:31-36``: H K H with the explicit centring matrix."""
    n = K.shape[0]
    H = torch.eye(n, dtype=K.dtype) - torch.ones(n, n, dtype=K.dtype) / n
    return H @ K @ H


def _rbf(X, sigma=None):
    """``gnndelete_nodeemb.py:38-46`` (with the ``math`` import the reference forgot)."""
    import math
    GX = X @ X.T
    KX = torch.diag(GX) - GX + (torch.diag(GX) - GX).T
    if sigma is None:
        sigma = math.sqrt(torch.median(KX[KX != 0]).item())
    return torch.exp(KX * (-0.5 / (sigma * sigma)))


def _cka(gram):
    """``gnndelete_nodeemb.py:48-66``: HSIC(X, Y) / sqrt(HSIC(X, X) HSIC(Y, Y))."""
    def hsic(X, Y):
        return torch.sum(_centering(gram(X)) * _centering(gram(Y)))
    return lambda X, Y: hsic(X, Y) / (torch.sqrt(hsic(X, X)) * torch.sqrt(hsic(Y, Y)))


def nodeemb_loss_fct(name):
    """``gnndelete_nodeemb.py:19-29, 69-92`` (``get_loss_fct``)."""
    if name == 'mse_mean':
        return torch.nn.MSELoss(reduction='mean')
    if name == 'mse_sum':
        return torch.nn.MSELoss(reduction='sum')
    if name in ('kld_mean', 'kld_sum'):
        red = 'batchmean' if name == 'kld_mean' else 'sum'
        return lambda logits, truth: 1 - torch.exp(-F.kl_div(F.log_softmax(logits, -1), truth.softmax(-1), reduction=red))
    if name == 'cosine_mean':
        return lambda logits, truth: (1 - F.cosine_similarity(logits, truth)).mean()
    if name == 'cosine_sum':
        return lambda logits, truth: (1 - F.cosine_similarity(logits, truth)).sum()
    if name == 'linear_cka':
        return _cka(lambda X: X @ X.T)
    if name == 'rbf_cka':
        return _cka(_rbf)
    raise NotImplementedError(name)


def nodeemb_epoch(model, data, neg_edge, z1_ori, z2_ori, optimizer, loss_type='both_layerwise', alpha=0.5,
                  loss_fct='mse_mean'):
    """One epoch of ``GNNDeleteNodeembTrainer.train_fullbatch``, ``gnndelete_nodeemb.py:191-299``, INCLUDING the
    backward / optimizer schedule of the chosen ``loss_type`` (the branches differ in which gradients are
    cleared: ``both_all`` never zeroes them, ``*_layerwise`` leave loss2's deletion1 gradient in ``.grad``).
    ``optimizer`` is the ``[optimizer1, optimizer2]`` pair for ``*layerwise`` types (delete_gnn.py:221-226), one
    Adam over both Del weights otherwise.  ``neg_edge`` is drawn once before the loop in the reference (:186-189)
    and supplied here.  Returns (loss, loss_r, loss_l) as logged (:301-307)."""
    fct = nodeemb_loss_fct(loss_fct)
    non_df = torch.ones(data.x.shape[0], dtype=torch.bool)                  # :171-175
    non_df[data.directed_df_edge_index.flatten().unique()] = False
    m1 = data.sdf_node_1hop_mask & non_df
    m2 = data.sdf_node_2hop_mask & non_df
    z1, z2 = model(data.x, data.train_pos_edge_index[:, data.sdf_mask], return_all_emb=True)     # :195
    pos_edge = data.train_pos_edge_index[:, data.df_mask]                   # :200
    embed1 = torch.cat([z1[pos_edge[0]], z1[pos_edge[1]]], dim=0)           # :203-207
    embed1_ori = torch.cat([z1_ori[neg_edge[0]], z1_ori[neg_edge[1]]], dim=0)
    embed2 = torch.cat([z2[pos_edge[0]], z2[pos_edge[1]]], dim=0)
    embed2_ori = torch.cat([z2_ori[neg_edge[0]], z2_ori[neg_edge[1]]], dim=0)
    loss_r1 = fct(embed1, embed1_ori)                                       # :209-210
    loss_r2 = fct(embed2, embed2_ori)
    loss_l1 = fct(z1[m1], z1_ori[m1])                                       # :213-214
    loss_l2 = fct(z2[m2], z2_ori[m2])
    if loss_type == 'both_all':                                             # :219-229
        loss_l, loss_r = loss_l1 + loss_l2, loss_r1 + loss_r2
        loss = alpha * loss_r + (1 - alpha) * loss_l
        loss.backward()
        optimizer.step()
    elif loss_type == 'both_layerwise':                                     # :231-246
        loss_l, loss_r = loss_l1 + loss_l2, loss_r1 + loss_r2
        loss1 = alpha * loss_r1 + (1 - alpha) * loss_l1
        loss1.backward(retain_graph=True)
        optimizer[0].step()
        optimizer[0].zero_grad()
        loss2 = alpha * loss_r2 + (1 - alpha) * loss_l2
        loss2.backward(retain_graph=True)
        optimizer[1].step()
        optimizer[1].zero_grad()
        loss = loss1 + loss2
    elif loss_type == 'only2_layerwise':                                    # :264-279
        loss_l, loss_r = loss_l1 + loss_l2, loss_r1 + loss_r2
        optimizer[0].zero_grad()
        loss2 = alpha * loss_r2 + (1 - alpha) * loss_l2
        loss2.backward()
        optimizer[1].step()
        optimizer[1].zero_grad()
        loss = loss2
    elif loss_type == 'only2_all':                                          # :281-289
        loss_l, loss_r = loss_l2, loss_r2
        loss = loss_l + alpha * loss_r
        loss.backward()
        optimizer.step()
        optimizer.zero_grad()
    elif loss_type == 'only1':                                              # :291-299
        loss_l, loss_r = loss_l1, loss_r1
        loss = loss_l + alpha * loss_r
        loss.backward()
        optimizer.step()
        optimizer.zero_grad()
    else:
        raise NotImplementedError(loss_type)
    return loss.detach(), loss_r.detach(), loss_l.detach()


def link_train_epoch(model, data, neg_edge_index, optimizer, retrain=False):
    """One epoch of ``Trainer.train_fullbatch`` (``framework/trainer/base.py:80-98``) or, with ``retrain``,
    of ``RetrainTrainer.train_fullbatch`` (``framework/trainer/retrain.py:56-72``: the same step on the
    retained edges): BCE link prediction on (edges, supplied negatives), backward to every parameter, step."""
    ei = data.train_pos_edge_index
    if retrain:
        ei = ei[:, data.dr_mask]
    z = model(data.x, ei)
    logits = model.decode(z, ei, neg_edge_index)
    label = torch.zeros(ei.shape[1] + neg_edge_index.shape[1], dtype=logits.dtype)     # get_link_labels, base.py:45-50
    label[:ei.shape[1]] = 1.
    loss = F.binary_cross_entropy_with_logits(logits, label)
    loss.backward()
    optimizer.step()
    optimizer.zero_grad()
    return loss.detach()


def split_edges(data, perm, val_ratio=0.05, test_ratio=0.1):
    """``train_test_split_edges_no_neg_adj_mask`` (``prepare_dataset.py:31-136``), homogeneous branch, with the
    permutation supplied: ``row < col`` edges, permuted, ``[test | val | train]`` (:45-50, :63-67, :99-100, :117-118).
    Returns (train, test, val) directed edge lists; the sampled negatives are random in the reference."""
    import math
    row, col = data.edge_index
    mask = row < col
    row, col = row[mask], col[mask]
    n_v = int(math.floor(val_ratio * row.size(0)))
    n_t = int(math.floor(test_ratio * row.size(0)))
    row, col = row[perm], col[perm]
    train = torch.stack([row[n_v + n_t:], col[n_v + n_t:]], dim=0)
    test = torch.stack([row[:n_t], col[:n_t]], dim=0)
    val = torch.stack([row[n_t:n_t + n_v], col[n_t:n_t + n_v]], dim=0)
    return train, test, val


def df_candidate_masks(train_pos_edge_index, test_pos_edge_index, num_nodes):
    """``prepare_dataset.py:203-215, 262-265``: edges inside / outside the 2-hop subgraph of the test edges."""
    _, _, _, mask = P.k_hop_subgraph(test_pos_edge_index.flatten().unique(), 2, train_pos_edge_index,
                                     num_nodes=num_nodes)
    return {'out': ~mask, 'in': mask}
