"""fp32 CPU oracle of the relation head (tube-pair transformer).

TEST INFRASTRUCTURE -- see oracle/__init__.py.  PINNED: every function here is
checked against outputs of the reference's own classes (imported from
/root/reference/models/relation_head by tests/golden/make_golden.py) in
tests/test_oracle_golden.py.

Functional restatement: takes the ``state_dict`` of the reference module.
"""
import math

import torch
import torch.nn.functional as F


def encoder_layer(sd, p, x, nhead):
    """torch ``nn.TransformerEncoderLayer`` (post-norm, ReLU, seq-first [S, B, E]) as
    instantiated at models/relation_head/base.py:32-35 and transformer.py:20-23."""
    s, b, e = x.shape
    hd = e // nhead
    qkv = F.linear(x, sd[p + 'self_attn.in_proj_weight'], sd[p + 'self_attn.in_proj_bias'])
    q, k, v = qkv.chunk(3, dim=-1)
    q = q.reshape(s, b * nhead, hd).transpose(0, 1)
    k = k.reshape(s, b * nhead, hd).transpose(0, 1)
    v = v.reshape(s, b * nhead, hd).transpose(0, 1)
    att = torch.softmax((q / math.sqrt(hd)) @ k.transpose(1, 2), dim=-1) @ v
    att = att.transpose(0, 1).reshape(s, b, e)
    att = F.linear(att, sd[p + 'self_attn.out_proj.weight'], sd[p + 'self_attn.out_proj.bias'])
    x = F.layer_norm(x + att, (e,), sd[p + 'norm1.weight'], sd[p + 'norm1.bias'], 1e-5)
    y = F.linear(F.relu(F.linear(x, sd[p + 'linear1.weight'], sd[p + 'linear1.bias'])),
                 sd[p + 'linear2.weight'], sd[p + 'linear2.bias'])
    return F.layer_norm(x + y, (e,), sd[p + 'norm2.weight'], sd[p + 'norm2.bias'], 1e-5)


def object_encoder(sd, x, num_heads=8, num_layers=2):
    """ObjectEncoder.forward, models/relation_head/base.py:26-40.

    x [N_tubes, T, 256] is consumed seq-first: sequence axis = tubes, batch axis = frames.
    """
    for l in range(num_layers):
        x = encoder_layer(sd, f'transformer_encoder.layers.{l}.', x, num_heads)
    return x


def pair_proposal(sd, encoded_subjects, encoded_objects):
    """PairProposalNetwork.forward, models/relation_head/base.py:49-62 (diagonal stays 0)."""
    sub = encoded_subjects.max(dim=1).values
    obj = encoded_objects.max(dim=1).values
    n = obj.shape[0]
    w1, b1 = sd['pair_ffn.0.weight'], sd['pair_ffn.0.bias']
    w2, b2 = sd['pair_ffn.2.weight'], sd['pair_ffn.2.bias']
    pair = torch.zeros(n, n)
    for i in range(n):
        comb = torch.cat([sub[i][None].expand(n, -1), obj], dim=-1)
        row = F.linear(F.relu(F.linear(comb, w1, b1)), w2, b2)[:, 0]
        row[i] = 0.0
        pair[i] = row
    return pair


def pick_top_pairs_eval(pred_matrix, num_total_pairs=100):
    """models/relation_head/test_utils.py:4-22."""
    n = pred_matrix.shape[0]
    m = pred_matrix.clone()
    m[torch.eye(n).bool()] = float('-inf')
    flat = m.view(-1)
    k = min(flat.shape[0], num_total_pairs)
    _, top = torch.topk(flat, k, sorted=True)
    return [[int(i // n), int(i % n)] for i in top.tolist() if i // n != i % n]


def concatenate_sub_obj(sub_feats, obj_feats, selected_pairs):
    """models/relation_head/train_utils.py:67-81."""
    return torch.stack([torch.cat([sub_feats[s], obj_feats[o]], dim=-1) for s, o in selected_pairs])


def positional_encoding(d_model, length):
    """PositionalEncoding buffer, models/relation_head/transformer.py:59-75."""
    position = torch.arange(length).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
    pe = torch.zeros(length, 1, d_model)
    pe[:, 0, 0::2] = torch.sin(position * div_term)
    pe[:, 0, 1::2] = torch.cos(position * div_term)
    return pe


def temporal_transformer(sd, x, num_layers=1):
    """TemporalTransformer.forward (eval), models/relation_head/transformer.py:35-56.

    x [P, T, 512] -> (span_pred [P, T, R], relation_pred [P, R]).
    """
    x = x.transpose(0, 1)
    x = x + sd['positional_encoding.pe'][:x.shape[0]]
    for l in range(num_layers):
        x = encoder_layer(sd, f'transformer_encoder.layers.{l}.', x, 4)
    x = F.layer_norm(x, (x.shape[-1],), sd['layer_norm.weight'], sd['layer_norm.bias'], 1e-5)
    x = x.transpose(0, 1)
    x = F.relu(F.linear(x, sd['fc1.weight'], sd['fc1.bias']))
    x = F.relu(F.linear(x, sd['fc2.weight'], sd['fc2.bias']))
    span_pred = F.linear(x, sd['span_head.weight'], sd['span_head.bias'])
    relation_pred = F.linear(x, sd['pred_head.weight'], sd['pred_head.bias']).max(dim=1).values
    return span_pred, relation_pred


def vanilla_model(sd, x):
    """VanillaModel.forward, models/relation_head/base.py:15-23."""
    x = F.relu(F.linear(x, sd['fc1.weight'], sd['fc1.bias']))
    x = F.relu(F.linear(x, sd['fc2.weight'], sd['fc2.bias']))
    span_pred = F.linear(x, sd['span_head.weight'], sd['span_head.bias'])
    relation_pred = F.linear(x, sd['pred_head.weight'], sd['pred_head.bias']).max(dim=1).values
    return span_pred, relation_pred


def handcrafted_filter(sd, x):
    """HandcraftedFilter.forward, models/relation_head/convolution.py:23-41: depthwise F.conv1d with the fixed taps
    [1/4, 1/2, 1, 1/2, 1/4] (padding 2) along time, then the VanillaModel heads."""
    c = x.shape[-1]
    w = torch.tensor([1 / 4, 1 / 2, 1, 1 / 2, 1 / 4], dtype=torch.float32).view(1, 1, -1).repeat(c, 1, 1)
    y = F.conv1d(x.permute(0, 2, 1), w, padding=2, groups=c).permute(0, 2, 1)
    return vanilla_model(sd, y)


def learnable_conv(sd, x, num_layers=1):
    """Learnable1DConv.forward, models/relation_head/convolution.py:63-75."""
    y = x.permute(0, 2, 1)
    for l in range(num_layers):
        w = sd[f'conv_layers.{2 * l}.weight']
        y = F.relu(F.conv1d(y, w, sd[f'conv_layers.{2 * l}.bias'], padding=w.shape[-1] // 2))
    return vanilla_model(sd, y.permute(0, 2, 1))


def generate_pairwise_results(span_pred, prob, selected_pairs):
    """models/relation_head/test_utils.py:56-84."""
    max_probs, max_indices = torch.max(prob, dim=1)
    _, order = torch.sort(max_probs, descending=True)
    results = []
    for p in order.tolist():
        r = int(max_indices[p])
        s, o = selected_pairs[p]
        results.append(dict(subject_index=s, object_index=o, relation=r,
                            relation_span=(span_pred[p, :, r].numpy() > 0).astype(float)))
    return results


def generate_results(span_pred, prob, selected_pairs):
    """models/relation_head/test_utils.py:25-53."""
    _, order = torch.sort(prob.flatten(), descending=True)
    nrel = prob.shape[1]
    results = []
    for idx in order.tolist():
        p, r = idx // nrel, idx % nrel
        s, o = selected_pairs[p]
        results.append(dict(subject_index=s, object_index=o, relation=r,
                            relation_span=(span_pred[p, :, r].numpy() > 0).astype(float)))
    return results


def relation_forward(sds, feats, num_top_pairs=100):
    """The forward section of tools/rel_test.py:35-67.

    sds: dict with the four state_dicts saved by tools/rel_train.py:223-231
    ('subject_encoder', 'object_encoder', 'pair_proposal_model', 'relation_model').
    """
    sub = object_encoder(sds['subject_encoder'], feats)
    obj = object_encoder(sds['object_encoder'], feats)
    pred_matrix = pair_proposal(sds['pair_proposal_model'], sub, obj)
    pairs = pick_top_pairs_eval(pred_matrix, num_top_pairs)
    cat = concatenate_sub_obj(sub, obj, pairs)
    span_pred, prob = temporal_transformer(sds['relation_model'], cat)
    return dict(sub=sub, obj=obj, pred_matrix=pred_matrix, pairs=pairs, span_pred=span_pred, prob=prob)
