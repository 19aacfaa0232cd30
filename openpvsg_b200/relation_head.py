"""B200 backend of the relation head (tube-pair transformer).

Same class names, constructor arguments and ``state_dict`` keys as
models/relation_head/{base,transformer,convolution,test_utils,train_utils}.py, so the
four-state_dict checkpoint written by tools/rel_train.py:223-231 loads unchanged and
tools/rel_test.py:39-67 runs against these classes.  Inference only.
"""
import math

import torch
import torch.nn as nn

from . import lib as _l
from . import ops


class _EncoderLayer(nn.Module):
    """Parameter container with torch ``nn.TransformerEncoderLayer`` key names; forward is
    post-norm / ReLU (the defaults the reference uses, base.py:32-35, transformer.py:20-23)."""

    def __init__(self, d_model, nhead, dim_feedforward):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(d_model, nhead)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.nhead = nhead

    @torch.no_grad()
    def forward_tokens(self, x, seq_axis, x_planes=None, want_planes=False):
        """x [A, B, E] contiguous; attention runs over axis ``seq_axis`` independently for every
        index of the other axis.  All row-wise layers work on the flat [A*B, E] view, and the
        attention kernel takes the (batch, sequence) structure through strides, so neither
        orientation needs a transpose.  x_planes: operand planes of x when the producer emitted them;
        want_planes: also return the planes of the result (-> (y, planes)).  The LayerNorms and the FFN
        hand planes to their consumers, so the only split pass left in a layer is the attention output."""
        A, Bx, E = x.shape
        a = self.self_attn
        if not x.is_contiguous():
            raise ops._l.PvsgError('encoder layer: contiguous [A,B,E] expected')
        o = torch.empty(A, Bx, E, device=x.device, dtype=torch.float32)
        tc = E // self.nhead in (32, 128)  # tensor-core attention on the projection's planes
        res = ops.linear(x_planes.view(A * Bx, E) if x_planes is not None else x.view(A * Bx, E), a.in_proj_weight, a.in_proj_bias,
                         out_mode='both' if tc else 'f32')
        qkv, planes = res if tc else (res, None)
        qkv = qkv.view(A, Bx, 3 * E)
        qv, ov = (qkv.permute(1, 0, 2), o.permute(1, 0, 2)) if seq_axis == 0 else (qkv, o)
        if planes is not None:
            ph, pl = planes.hi.view(A, Bx, 3 * E), planes.lo.view(A, Bx, 3 * E)
            if seq_axis == 0:
                ph, pl = ph.permute(1, 0, 2), pl.permute(1, 0, 2)
            ops.attention(qv[..., :E], ops.Split(ph[..., E:2 * E], pl[..., E:2 * E]),
                          ops.Split(ph[..., 2 * E:], pl[..., 2 * E:]), self.nhead, out=ov)
        else:
            ops.attention(qv[..., :E], qv[..., E:2 * E], qv[..., 2 * E:], self.nhead, out=ov)
        y = ops.linear(o, a.out_proj.weight, a.out_proj.bias, residual=x)
        y, ys = ops.layernorm(y, self.norm1.weight, self.norm1.bias, self.norm1.eps, out_split=True)
        h = ops.linear(ys if ys is not None else y, self.linear1.weight, self.linear1.bias, act=ops.ACT_RELU, out_mode='split')
        z = ops.linear(h, self.linear2.weight, self.linear2.bias, residual=y)
        z, zs = ops.layernorm(z, self.norm2.weight, self.norm2.bias, self.norm2.eps, out_split=True)
        return (z, zs) if want_planes else z


class _Encoder(nn.Module):
    def __init__(self, d_model, nhead, dim_feedforward, num_layers):
        super().__init__()
        self.layers = nn.ModuleList([_EncoderLayer(d_model, nhead, dim_feedforward) for _ in range(num_layers)])


class _Inference(nn.Module):
    def train(self, mode=True):
        if mode:
            raise NotImplementedError('openpvsg_b200.relation_head implements inference only')
        return super().train(False)


class ObjectEncoder(_Inference):
    """base.py:26-40.  forward(x [N_tubes, T, 256]): the reference feeds this seq-first, i.e.
    the SEQUENCE axis is the tubes and the batch axis is the frames."""

    def __init__(self, feature_dim=256, hidden_dim=512, num_heads=8, num_layers=2):
        super().__init__()
        self.transformer_encoder = _Encoder(feature_dim, num_heads, hidden_dim, num_layers)
        self.eval()

    @torch.no_grad()
    def forward(self, x):
        y, ys = x.contiguous(), None
        for layer in self.transformer_encoder.layers:
            y, ys = layer.forward_tokens(y, 0, ys, want_planes=True)   # sequence = tubes, batch = frames
        return y


class PairProposalNetwork(_Inference):
    """base.py:43-62, evaluated in factorised form: W1 [hidden, 2F] splits into a subject half and
    an object half, so the N^2 MLP evaluations reduce to two [N, hidden] GEMMs and one
    N^2 x hidden relu-dot kernel (pvsg_pair_proposal).  The result stays on the device (the
    reference allocates it on the CPU, base.py:53)."""

    def __init__(self, feature_dim, hidden_dim):
        super().__init__()
        self.pair_ffn = nn.Sequential(nn.Linear(feature_dim * 2, hidden_dim), nn.ReLU(), nn.Linear(hidden_dim, 1))
        self.feature_dim = feature_dim
        self.eval()

    @torch.no_grad()
    def forward(self, encoded_subjects, encoded_objects):
        F_ = self.feature_dim
        sub = ops.max_over_time(encoded_subjects)
        obj = ops.max_over_time(encoded_objects)
        w1, b1 = self.pair_ffn[0].weight, self.pair_ffn[0].bias
        U = ops.linear(sub, w1[:, :F_], b1)
        V = ops.linear(obj, w1[:, F_:])
        return ops.pair_proposal(U, V, self.pair_ffn[2].weight.view(-1), self.pair_ffn[2].bias)


class PositionalEncoding(nn.Module):
    """transformer.py:59-81 (buffer ``pe`` [max_len, 1, d_model])."""

    def __init__(self, d_model, dropout=0.1, max_len=5000):
        super().__init__()
        position = torch.arange(max_len).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
        pe = torch.zeros(max_len, 1, d_model)
        pe[:, 0, 0::2] = torch.sin(position * div_term)
        pe[:, 0, 1::2] = torch.cos(position * div_term)
        self.register_buffer('pe', pe)


def _heads(mod, x):
    """fc1/fc2/span_head/pred_head tail shared by all relation models (base.py:16-23)."""
    P, T, _ = x.shape
    h = ops.linear(x, mod.fc1.weight, mod.fc1.bias, act=ops.ACT_RELU, out_mode='split')     # hidden layers as planes only
    h = ops.linear(h, mod.fc2.weight, mod.fc2.bias, act=ops.ACT_RELU, out_mode='split')
    span_pred = ops.linear(h, mod.span_head.weight, mod.span_head.bias)
    rel = ops.linear(h, mod.pred_head.weight, mod.pred_head.bias)
    return span_pred, ops.max_over_time(rel)


class TemporalTransformer(_Inference):
    """transformer.py:8-56.  forward(x [P, T, 512]) -> (span_pred [P,T,R], relation_pred [P,R]).
    ``forward_pairs`` fuses concatenate_sub_obj + the positional-encoding add into one gather."""

    def __init__(self, input_dim=512, num_relations=57, num_transformer_layers=1, dropout_rate=0.1):
        super().__init__()
        self.num_relations = num_relations
        self.positional_encoding = PositionalEncoding(input_dim, dropout=dropout_rate)
        self.transformer_encoder = _Encoder(input_dim, 4, 512, num_transformer_layers)
        self.layer_norm = nn.LayerNorm(input_dim)
        self.fc1 = nn.Linear(input_dim, input_dim // 2)
        self.fc2 = nn.Linear(input_dim // 2, input_dim // 4)
        self.span_head = nn.Linear(input_dim // 4, num_relations)
        self.pred_head = nn.Linear(input_dim // 4, num_relations)
        self.eval()

    @torch.no_grad()
    def _encode(self, x):
        xs = None
        for layer in self.transformer_encoder.layers:
            x, xs = layer.forward_tokens(x, 1, xs, want_planes=True)   # sequence = frames, batch = pairs
        x, xs = ops.layernorm(x, self.layer_norm.weight, self.layer_norm.bias, self.layer_norm.eps, out_split=True)
        return _heads(self, xs if xs is not None else x)

    @torch.no_grad()
    def forward(self, x):
        P, T, E = x.shape
        pe = self.positional_encoding.pe[:T, 0]                       # [T,E]
        x = ops.add_rowvec(x.contiguous().view(P, T * E), pe.reshape(-1)).view(P, T, E)
        return self._encode(x)

    @torch.no_grad()
    def forward_pairs(self, sub_feats, obj_feats, pairs):
        """pairs int32 [P,2] on the device."""
        pe = self.positional_encoding.pe.view(self.positional_encoding.pe.shape[0], -1)
        return self._encode(ops.gather_pairs(sub_feats, obj_feats, pairs, pe))


class VanillaModel(_Inference):
    """base.py:6-23."""

    def __init__(self, input_dim, num_relations):
        super().__init__()
        self.fc1 = nn.Linear(input_dim, input_dim // 2)
        self.fc2 = nn.Linear(input_dim // 2, input_dim // 4)
        self.span_head = nn.Linear(input_dim // 4, num_relations)
        self.pred_head = nn.Linear(input_dim // 4, num_relations)
        self.eval()

    @torch.no_grad()
    def forward(self, x):
        return _heads(self, x.contiguous())


class HandcraftedFilter(_Inference):
    """convolution.py:6-41: a fixed 5-tap temporal filter [1/4, 1/2, 1, 1/2, 1/4] on every channel (depthwise
    F.conv1d, padding 2), then the VanillaModel heads.  ``tools/rel_test.py:167-175`` selects it with
    ``--model-name filter``."""

    def __init__(self, feat_dim, num_relations):
        super().__init__()
        self.num_relations = num_relations
        self.filter_weights = torch.tensor([1 / 4, 1 / 2, 1, 1 / 2, 1 / 4], dtype=torch.float32)   # plain attribute, as the reference
        self.expanded_filter_weights = self.filter_weights.view(1, 1, -1).repeat(feat_dim, 1, 1)
        self.fc1 = nn.Linear(feat_dim, feat_dim // 2)
        self.fc2 = nn.Linear(feat_dim // 2, feat_dim // 4)
        self.span_head = nn.Linear(feat_dim // 4, num_relations)
        self.pred_head = nn.Linear(feat_dim // 4, num_relations)
        self.eval()

    @torch.no_grad()
    def forward(self, x):
        return _heads(self, ops.temporal_fir(x.contiguous(), self.filter_weights.to(x.device)))


class Learnable1DConv(_Inference):
    """convolution.py:44-75: ``num_layers`` x (Conv1d(C, C, k, padding k // 2) + ReLU) along time, then the
    VanillaModel heads.  Each Conv1d is ONE GEMM over the k temporal taps laid side by side
    (``ops.temporal_unfold`` -> [P*T, k*C] x [C, k*C]^T, bias + ReLU in the epilogue)."""

    def __init__(self, input_dim, num_relations, kernel_size=5, num_layers=1):
        super().__init__()
        self.num_relations = num_relations
        layers = []
        for _ in range(num_layers):
            layers += [nn.Conv1d(input_dim, input_dim, kernel_size, padding=kernel_size // 2), nn.ReLU()]
        self.conv_layers = nn.Sequential(*layers)
        self.fc1 = nn.Linear(input_dim, input_dim // 2)
        self.fc2 = nn.Linear(input_dim // 2, input_dim // 4)
        self.span_head = nn.Linear(input_dim // 4, num_relations)
        self.pred_head = nn.Linear(input_dim // 4, num_relations)
        self._w = {}
        self.eval()

    def _tap_major(self, conv):
        """Conv1d weight [Cout, Cin, k] -> [Cout, k*Cin] matching temporal_unfold's column order."""
        key = (id(conv), conv.weight._version, conv.weight.data_ptr())
        if key not in self._w:
            self._w = {key: conv.weight.permute(0, 2, 1).reshape(conv.out_channels, -1).contiguous()}
        return self._w[key]

    @torch.no_grad()
    def forward(self, x):
        x = x.contiguous()
        for conv in self.conv_layers:
            if isinstance(conv, nn.Conv1d):
                if conv.kernel_size[0] % 2 == 0 or conv.padding[0] != conv.kernel_size[0] // 2:
                    raise NotImplementedError('Learnable1DConv: odd kernel with "same" padding (the reference form)')
                x = ops.linear(ops.temporal_unfold(x, conv.kernel_size[0]), self._tap_major(conv), conv.bias, act=ops.ACT_RELU)
        return _heads(self, x)


# --------------------------------------------------------------------------------------
# test_utils.py / train_utils.py helpers
# --------------------------------------------------------------------------------------
@torch.no_grad()
def pick_top_pairs_device(pred_matrix, num_total_pairs=100):
    """Device form of pick_top_pairs_eval: (pairs int32 [k,2], count int32 [1]), no host sync."""
    return ops.top_pairs(pred_matrix, num_total_pairs)


@torch.no_grad()
def pick_top_pairs_eval(pred_matrix, num_total_pairs=100):
    """test_utils.py:4-22 -> python list of [s, o]."""
    pairs, n = pick_top_pairs_device(pred_matrix, num_total_pairs)
    return pairs[:int(n.item())].cpu().tolist()


@torch.no_grad()
def concatenate_sub_obj(sub_feats, obj_feats, selected_pairs):
    """train_utils.py:67-81."""
    pairs = selected_pairs if torch.is_tensor(selected_pairs) else \
        torch.tensor(selected_pairs, dtype=torch.int32, device=sub_feats.device).view(-1, 2)
    return ops.gather_pairs(sub_feats, obj_feats, pairs.to(torch.int32))


def _results(order, rel_of, span_pred, selected_pairs):
    span = (span_pred > 0).cpu().numpy()  # one D2H copy instead of one per result
    out = []
    for p, r in zip(order, rel_of):
        s, o = selected_pairs[p]
        out.append(dict(subject_index=s, object_index=o, relation=int(r),
                        relation_span=span[p, :, r].astype(float)))
    return out


@torch.no_grad()
def generate_pairwise_results(span_pred, prob, selected_pairs):
    """test_utils.py:56-84."""
    prob_c = prob.cpu()
    max_probs, max_indices = torch.max(prob_c, dim=1)
    _, order = torch.sort(max_probs, descending=True)
    order = order.tolist()
    return _results(order, [int(max_indices[p]) for p in order], span_pred, selected_pairs)


@torch.no_grad()
def generate_results(span_pred, prob, selected_pairs):
    """test_utils.py:25-53."""
    prob_c = prob.cpu()
    _, order = torch.sort(prob_c.flatten(), descending=True)
    nrel = prob_c.shape[1]
    order = order.tolist()
    return _results([i // nrel for i in order], [i % nrel for i in order], span_pred, selected_pairs)


@torch.no_grad()
def relation_forward(subject_encoder, object_encoder, pair_proposal_model, relation_model, feats,
                     num_top_pairs=100, graph=False):
    """The forward section of tools/rel_test.py:35-67 without host round trips between stages.
    graph=True: the ~75 launches are captured once per (N, T, num_top_pairs) into a CUDA graph and replayed (the
    stage is latency-bound: ~25 us of work per launch); the returned tensors are then views of the graph's static
    outputs, valid until the next call with the same shape."""
    if graph:
        return _graphed_forward((subject_encoder, object_encoder, pair_proposal_model, relation_model), feats, num_top_pairs)
    sub = subject_encoder(feats)
    obj = object_encoder(feats)
    pred_matrix = pair_proposal_model(sub, obj)
    pairs, _n = pick_top_pairs_device(pred_matrix, num_top_pairs)
    # test_utils.py:4-22 takes topk(min(N^2, k)) of the matrix with its diagonal at -inf: the count is known on the host
    pairs = pairs[:min(int(num_top_pairs), feats.shape[0] * feats.shape[0])]
    span_pred, prob = relation_model.forward_pairs(sub, obj, pairs) if hasattr(relation_model, 'forward_pairs') \
        else relation_model(concatenate_sub_obj(sub, obj, pairs))
    return dict(sub=sub, obj=obj, pred_matrix=pred_matrix, pairs=pairs, span_pred=span_pred, prob=prob)


def _graphed_forward(models, feats, num_top_pairs):
    holder = models[3]
    epoch = hash(tuple(p._version for m in models for p in m.parameters()) + tuple(p.data_ptr() for m in models for p in m.parameters()))
    key = (tuple(feats.shape), int(num_top_pairs), str(feats.device), tuple(id(m) for m in models), epoch)
    cache = holder.__dict__.setdefault('_pvsg_graphs', {})
    hit = cache.get(key)
    if hit is None:
        _l.handle(feats.device.index if feats.device.index is not None else torch.cuda.current_device())
        cache.clear()                      # one shape at a time: a changed clip length / weights epoch drops the old graph
        static_in = torch.empty_like(feats)
        static_in.copy_(feats)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):             # warm-up: weight-plane caches are filled outside the capture
                relation_forward(*models, static_in, num_top_pairs)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = relation_forward(*models, static_in, num_top_pairs)
        hit = cache[key] = (g, static_in, out)
    g, static_in, out = hit
    static_in.copy_(feats, non_blocking=True)
    g.replay()
    return out
