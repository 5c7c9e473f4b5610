"""Drop-in torch.nn.Module mirrors of the reference's hot-path layers.

Same class names, constructor signatures, child-module names (hence state_dict keys),
initialisation order and forward signatures as ``fuxictr.pytorch.layers`` — the modules
are *parameter containers*; their forwards dispatch to the sm_100a kernels of
libfuxictr_b200.so through fuxictr_b200.functional.  CUDA tensors are required: a CPU
tensor raises (there is no CPU implementation on this path).

Reference files (relative to the reference root):
  FeatureEmbedding / FeatureEmbeddingDict  fuxictr/pytorch/layers/embeddings/feature_embedding.py:30-297
  MaskedAveragePooling / MaskedSumPooling  fuxictr/pytorch/layers/pooling.py:23-73
  LogisticRegression                       fuxictr/pytorch/layers/blocks/logistic_regression.py:24-59
  FactorizationMachine                     fuxictr/pytorch/layers/blocks/factorization_machine.py:25-59
  InnerProductInteraction                  fuxictr/pytorch/layers/interactions/inner_product.py:23-70
  CrossInteraction / CrossNet / CrossNetV2 fuxictr/pytorch/layers/interactions/cross_net.py:24-129
  CompressedInteractionNet                 fuxictr/pytorch/layers/interactions/compressed_interaction_net.py:23-76
  DIN_Attention                            fuxictr/pytorch/layers/attentions/target_attention.py:26-92
  Dice                                     fuxictr/pytorch/layers/activations.py:24-51
  MLP_Block                                fuxictr/pytorch/layers/blocks/mlp_block.py:24-96
"""
import sys
from collections import OrderedDict
from functools import partial  # noqa: F401  (initializer strings use it)

import numpy as np  # noqa: F401
import torch
from torch import nn

from . import _lib
from . import functional as F2
from ._lib import (B2_POOL_NONE, B2_POOL_SUM, B2_POOL_MEAN, B2_ACT_NONE, B2_ACT_RELU,
                   B2_ACT_SIGMOID, FM_PRODUCT_SUM, FM_BI_INTERACTION, FM_INNER_PRODUCT)

layers = sys.modules[__name__]  # so feature_encoder strings like "layers.MaskedSumPooling()" resolve


def not_in_whitelist(element, whitelist=[]):
    """fuxictr/utils.py: whitelist test used by the feature filters."""
    if not whitelist:
        return False
    if not isinstance(whitelist, list):
        whitelist = [whitelist]
    return element not in whitelist


def get_initializer(initializer):
    """torch_utils.py:175-194: initializer strings are evaluated."""
    if isinstance(initializer, str):
        try:
            initializer = eval(initializer)
        except Exception:
            raise ValueError("initializer={} is not supported.".format(initializer))
    return initializer


def get_activation(activation, hidden_units=None):
    """torch_utils.py:137-173."""
    if isinstance(activation, str):
        if activation.lower() in ["prelu", "dice"]:
            assert type(hidden_units) == int
        if activation.lower() == "relu":
            return nn.ReLU()
        elif activation.lower() == "sigmoid":
            return nn.Sigmoid()
        elif activation.lower() == "tanh":
            return nn.Tanh()
        elif activation.lower() == "softmax":
            return nn.Softmax(dim=-1)
        elif activation.lower() == "prelu":
            return nn.PReLU(hidden_units, init=0.1)
        elif activation.lower() == "dice":
            return Dice(hidden_units)
        else:
            return getattr(nn, activation)()
    elif isinstance(activation, list):
        if hidden_units is not None:
            assert len(activation) == len(hidden_units)
            return [get_activation(act, units) for act, units in zip(activation, hidden_units)]
        else:
            return [get_activation(act) for act in activation]
    return activation


# --------------------------------------------------------------------------------------
# Pooling encoders (fused into the gather when used as a feature_encoder)
# --------------------------------------------------------------------------------------
class MaskedAveragePooling(nn.Module):
    def __init__(self):
        super(MaskedAveragePooling, self).__init__()

    def forward(self, embedding_matrix, mask=None):
        # stand-alone use on an already materialised (B, L, D) tensor: glue ops
        sum_out = torch.sum(embedding_matrix, dim=1)
        if mask is None:
            mask = embedding_matrix.sum(dim=-1) != 0
        return sum_out / (mask.float().sum(-1, keepdim=True) + 1e-12)


class MaskedSumPooling(nn.Module):
    def __init__(self):
        super(MaskedSumPooling, self).__init__()

    def forward(self, embedding_matrix):
        return torch.sum(embedding_matrix, dim=1)


# --------------------------------------------------------------------------------------
# Embeddings
# --------------------------------------------------------------------------------------
class FeatureEmbeddingDict(nn.Module):
    def __init__(self,
                 feature_map,
                 embedding_dim,
                 embedding_initializer="partial(nn.init.normal_, std=1e-4)",
                 required_feature_columns=None,
                 not_required_feature_columns=None,
                 use_pretrain=True,
                 use_sharing=True):
        super(FeatureEmbeddingDict, self).__init__()
        self._feature_map = feature_map
        self.required_feature_columns = required_feature_columns
        self.not_required_feature_columns = not_required_feature_columns
        self.use_pretrain = use_pretrain
        self.embedding_initializer = get_initializer(embedding_initializer)
        self.embedding_layers = nn.ModuleDict()
        self.feature_encoders = nn.ModuleDict()
        self._plans = {}
        for feature, feature_spec in self._feature_map.features.items():
            if self.is_required(feature):
                if not (use_pretrain and use_sharing) and embedding_dim == 1:
                    feat_dim = 1  # in case for LR
                    if feature_spec["type"] == "sequence":
                        self.feature_encoders[feature] = MaskedSumPooling()
                else:
                    feat_dim = feature_spec.get("embedding_dim", embedding_dim)
                    if feature_spec.get("feature_encoder", None):
                        self.feature_encoders[feature] = self.get_feature_encoder(feature_spec["feature_encoder"])
                    else:
                        if feature_spec["type"] == "embedding":
                            pretrain_dim = feature_spec.get("pretrain_dim", feat_dim)
                            self.feature_encoders[feature] = nn.Linear(pretrain_dim, feat_dim, bias=False)

                if use_sharing and feature_spec.get("share_embedding") in self.embedding_layers:
                    self.embedding_layers[feature] = self.embedding_layers[feature_spec["share_embedding"]]
                    continue

                if feature_spec["type"] == "numeric":
                    self.embedding_layers[feature] = nn.Linear(1, feat_dim, bias=False)
                elif feature_spec["type"] in ["categorical", "sequence"]:
                    if use_pretrain and "pretrained_emb" in feature_spec:
                        raise NotImplementedError(
                            "feature %s: pretrained_emb is outside the B200 hot path "
                            "(SURVEY.md section 2 row 2); keep the reference module for it" % feature)
                    padding_idx = feature_spec.get("padding_idx", None)
                    self.embedding_layers[feature] = nn.Embedding(feature_spec["vocab_size"],
                                                                  feat_dim,
                                                                  padding_idx=padding_idx)
                elif feature_spec["type"] == "embedding":
                    self.embedding_layers[feature] = nn.Identity()
        self.init_weights()

    def get_feature_encoder(self, encoder):
        try:
            if type(encoder) == list:
                encoder_layer = nn.Sequential(*[eval(enc) for enc in encoder])
            else:
                encoder_layer = eval(encoder)
            return encoder_layer
        except Exception:
            raise ValueError("feature_encoder={} is not supported.".format(encoder))

    def init_weights(self):
        for k, v in self.embedding_layers.items():
            if "share_embedding" in self._feature_map.features[k]:
                continue
            if type(v) == nn.Embedding:
                if v.padding_idx is not None:
                    self.embedding_initializer(v.weight[1:, :])  # set padding_idx to zero
                else:
                    self.embedding_initializer(v.weight)

    def is_required(self, feature):
        feature_spec = self._feature_map.features[feature]
        if feature_spec["type"] == "meta":
            return False
        elif self.required_feature_columns and (feature not in self.required_feature_columns):
            return False
        elif self.not_required_feature_columns and (feature in self.not_required_feature_columns):
            return False
        else:
            return True

    def dict2tensor(self, embedding_dict, flatten_emb=False, feature_list=[], feature_source=[],
                    feature_type=[]):
        feature_emb_list = []
        for feature, feature_spec in self._feature_map.features.items():
            if feature_list and not_in_whitelist(feature, feature_list):
                continue
            if feature_source and not_in_whitelist(feature_spec["source"], feature_source):
                continue
            if feature_type and not_in_whitelist(feature_spec["type"], feature_type):
                continue
            if feature in embedding_dict:
                feature_emb_list.append(embedding_dict[feature])
        if flatten_emb:
            feature_emb = torch.cat(feature_emb_list, dim=-1)
        else:
            feature_emb = torch.stack(feature_emb_list, dim=1)
        return feature_emb

    # ---- fused path -------------------------------------------------------------------
    def _active_features(self, inputs, feature_source, feature_type):
        names = []
        for feature in inputs.keys():
            feature_spec = self._feature_map.features[feature]
            if feature_source and not_in_whitelist(feature_spec["source"], feature_source):
                continue
            if feature_type and not_in_whitelist(feature_spec["type"], feature_type):
                continue
            if feature in self.embedding_layers:
                names.append(feature)
        return names

    def _is_fusable(self, feature):
        """Plain nn.Embedding lookup, optionally followed by a Masked{Sum,Average}Pooling."""
        spec = self._feature_map.features[feature]
        if spec["type"] not in ("categorical", "sequence"):
            return False
        if type(self.embedding_layers[feature]) != nn.Embedding:
            return False
        if feature in self.feature_encoders:
            enc = self.feature_encoders[feature]
            # by name: the reference's own pooling classes qualify too (fuxictr_b200.patch)
            if not (spec["type"] == "sequence" and
                    type(enc).__name__ in ("MaskedSumPooling", "MaskedAveragePooling")):
                return False
        return True

    def _plan(self, names, order):
        """Build (and cache) the launch plan for the fusable features `names` laid out in `order`."""
        key = (tuple(names), tuple(order))
        plan = self._plans.get(key)
        if plan is not None:
            return plan
        tables, slot_of, fields = [], {}, []
        for feature in order:
            spec = self._feature_map.features[feature]
            emb = self.embedding_layers[feature]
            if id(emb) not in slot_of:
                slot_of[id(emb)] = len(tables)
                tables.append(emb)
            seq_len = spec["max_len"] if spec["type"] == "sequence" else 1
            pool = B2_POOL_NONE
            if feature in self.feature_encoders:
                pool = B2_POOL_SUM if type(self.feature_encoders[feature]).__name__ == "MaskedSumPooling" \
                    else B2_POOL_MEAN
            fields.append(F2.GatherField(feature, slot_of[id(emb)], emb.embedding_dim, seq_len, pool,
                                         emb.padding_idx))
        plan = (F2.GatherPlan(fields), tables)
        self._plans[key] = plan
        return plan

    def _fused_arena(self, inputs, order):
        plan, tables = self._plan(order, order)
        idx = [inputs[f] for f in order]
        arena = F2.embed_gather(plan, idx, [t.weight for t in tables])
        return plan, arena

    def forward(self, inputs, feature_source=[], feature_type=[]):
        names = self._active_features(inputs, feature_source, feature_type)
        feature_emb_dict = OrderedDict()
        fusable = [f for f in names if self._is_fusable(f)]
        fused_out = {}
        if fusable:
            plan, arena = self._fused_arena(inputs, fusable)
            B = arena.shape[0]
            parts = arena.split(plan.widths, dim=1) if len(fusable) > 1 else (arena,)
            for field, part in zip(plan.fields, parts):
                if field.seq_len > 1 and field.pool == B2_POOL_NONE:
                    part = part.reshape(B, field.seq_len, field.dim)
                fused_out[field.name] = part
        for feature in names:
            if feature in fused_out:
                feature_emb_dict[feature] = fused_out[feature]
                continue
            # features outside the fused kernel (numeric / embedding-type / custom encoders):
            # the reference's own per-feature ops (feature_embedding.py:279-295)
            feature_spec = self._feature_map.features[feature]
            if feature_spec["type"] == "numeric":
                embeddings = self.embedding_layers[feature](inputs[feature].float().view(-1, 1))
            elif feature_spec["type"] in ("categorical", "sequence"):
                embeddings = self.embedding_layers[feature](inputs[feature].long())
            elif feature_spec["type"] == "embedding":
                embeddings = self.embedding_layers[feature](inputs[feature].float())
            else:
                raise NotImplementedError
            if feature in self.feature_encoders:
                embeddings = self.feature_encoders[feature](embeddings)
            feature_emb_dict[feature] = embeddings
        return feature_emb_dict

    def forward_tensor(self, inputs, feature_source=[], feature_type=[], flatten_emb=False):
        """FeatureEmbedding.forward in one launch: the gather writes the stacked (B,F,D) /
        concatenated (B, sum D) tensor directly (no dict, no torch.stack/cat)."""
        names = self._active_features(inputs, feature_source, feature_type)
        order = [f for f in self._feature_map.features.keys() if f in set(names)]
        if order and all(self._is_fusable(f) for f in order):
            plan, arena = self._fused_arena(inputs, order)
            unpooled = any(f.seq_len > 1 and f.pool == B2_POOL_NONE for f in plan.fields)
            if flatten_emb and not unpooled:
                return arena
            dims = set(f.dim for f in plan.fields)
            if not flatten_emb and not unpooled and len(dims) == 1:
                return arena.view(arena.shape[0], len(plan.fields), plan.fields[0].dim)
        feature_emb_dict = self.forward(inputs, feature_source=feature_source, feature_type=feature_type)
        return self.dict2tensor(feature_emb_dict, flatten_emb=flatten_emb)


def fused_front(embedding_layer, lr_layer, X, want_fm):
    """One launch for FeatureEmbedding + (FM product_sum) + (LogisticRegression): returns
    (feature_emb (B,F,D), logit (B,1)), or None when the configuration needs the general path
    (sequence / numeric features, mixed dims, dim % 4 != 0)."""
    fed = embedding_layer.embedding_layer
    names = fed._active_features(X, [], [])
    order = [f for f in fed._feature_map.features.keys() if f in set(names)]
    if not order or not all(fed._is_fusable(f) for f in order):
        return None
    plan, tables = fed._plan(order, order)
    lr_plan = lr_tables = bias = None
    if lr_layer is not None:
        lfed = lr_layer.embedding_layer.embedding_layer
        lnames = lfed._active_features(X, [], [])
        lorder = [f for f in lfed._feature_map.features.keys() if f in set(lnames)]
        if lorder != order or not all(lfed._is_fusable(f) for f in lorder):
            return None
        lr_plan, lr_tables = lfed._plan(lorder, lorder)
        bias = lr_layer.bias
    if not F2.front_supported(plan, lr_plan):
        return None
    arena, logit = F2.front(plan, lr_plan, [X[f] for f in order], [t.weight for t in tables],
                            [t.weight for t in lr_tables] if lr_tables else [], bias, want_fm)
    return arena.view(arena.shape[0], len(plan.fields), plan.fields[0].dim), logit


class FeatureEmbedding(nn.Module):
    def __init__(self,
                 feature_map,
                 embedding_dim,
                 embedding_initializer="partial(nn.init.normal_, std=1e-4)",
                 required_feature_columns=None,
                 not_required_feature_columns=None,
                 use_pretrain=True,
                 use_sharing=True):
        super(FeatureEmbedding, self).__init__()
        self.embedding_layer = FeatureEmbeddingDict(feature_map,
                                                    embedding_dim,
                                                    embedding_initializer=embedding_initializer,
                                                    required_feature_columns=required_feature_columns,
                                                    not_required_feature_columns=not_required_feature_columns,
                                                    use_pretrain=use_pretrain,
                                                    use_sharing=use_sharing)

    def forward(self, X, feature_source=[], feature_type=[], flatten_emb=False):
        return self.embedding_layer.forward_tensor(X, feature_source=feature_source,
                                                   feature_type=feature_type, flatten_emb=flatten_emb)


# --------------------------------------------------------------------------------------
# LR / FM
# --------------------------------------------------------------------------------------
class LogisticRegression(nn.Module):
    def __init__(self, feature_map, use_bias=True):
        super(LogisticRegression, self).__init__()
        self.bias = nn.Parameter(torch.zeros(1), requires_grad=True) if use_bias else None
        # A trick for quick one-hot encoding in LR
        self.embedding_layer = FeatureEmbedding(feature_map, 1, use_pretrain=False, use_sharing=False)
        self._lr_plans = {}

    def forward(self, X):
        fed = self.embedding_layer.embedding_layer
        names = fed._active_features(X, [], [])
        order = [f for f in fed._feature_map.features.keys() if f in set(names)]
        if order and all(fed._is_fusable(f) for f in order):
            key = tuple(order)
            if key not in self._lr_plans:
                self._lr_plans[key] = fed._plan(order, order)
            plan, tables = self._lr_plans[key]
            return F2.lr_forward(plan, [X[f] for f in order], [t.weight for t in tables], self.bias)
        embed_weights = self.embedding_layer(X)
        output = embed_weights.sum(dim=1)
        if self.bias is not None:
            output = output + self.bias
        return output


class InnerProductInteraction(nn.Module):
    """ output: product_sum (bs x 1),
                bi_interaction (bs * dim),
                inner_product (bs x f^2/2),
                elementwise_product (bs x f^2/2 x emb_dim)
    """
    def __init__(self, num_fields, output="product_sum"):
        super(InnerProductInteraction, self).__init__()
        self._output_type = output
        if output not in ["product_sum", "bi_interaction", "inner_product", "elementwise_product"]:
            raise ValueError("InnerProductInteraction output={} is not supported.".format(output))
        if output == "inner_product":
            self.interaction_units = int(num_fields * (num_fields - 1) / 2)
            self.triu_mask = nn.Parameter(torch.triu(torch.ones(num_fields, num_fields), 1).bool(),
                                          requires_grad=False)
        elif output == "elementwise_product":
            self.triu_index = nn.Parameter(torch.triu_indices(num_fields, num_fields, offset=1),
                                           requires_grad=False)

    def forward(self, feature_emb):
        if self._output_type == "product_sum":
            return F2.fm_interaction(feature_emb, FM_PRODUCT_SUM)
        elif self._output_type == "bi_interaction":
            return F2.fm_interaction(feature_emb, FM_BI_INTERACTION)
        elif self._output_type == "inner_product":
            return F2.fm_interaction(feature_emb, FM_INNER_PRODUCT)
        else:  # elementwise_product (PNN family, outside the five in-scope models): glue ops
            emb1 = torch.index_select(feature_emb, 1, self.triu_index[0])
            emb2 = torch.index_select(feature_emb, 1, self.triu_index[1])
            return emb1 * emb2


class FactorizationMachine(nn.Module):
    def __init__(self, feature_map):
        super(FactorizationMachine, self).__init__()
        self.fm_layer = InnerProductInteraction(feature_map.num_fields, output="product_sum")
        self.lr_layer = LogisticRegression(feature_map, use_bias=True)

    def forward(self, X, feature_emb):
        lr_out = self.lr_layer(X)
        fm_out = self.fm_layer(feature_emb)
        return fm_out + lr_out


# --------------------------------------------------------------------------------------
# Cross networks
# --------------------------------------------------------------------------------------
class CrossInteraction(nn.Module):
    def __init__(self, input_dim):
        super(CrossInteraction, self).__init__()
        self.weight = nn.Linear(input_dim, 1, bias=False)
        self.bias = nn.Parameter(torch.zeros(input_dim))

    def forward(self, X_0, X_i):
        # stand-alone use (CrossNet fuses all its CrossInteraction layers into one launch)
        return F2.linear_act(X_i, self.weight.weight, None, B2_ACT_NONE) * X_0 + self.bias


class CrossNet(nn.Module):
    def __init__(self, input_dim, num_layers):
        super(CrossNet, self).__init__()
        self.num_layers = num_layers
        self.cross_net = nn.ModuleList(CrossInteraction(input_dim)
                                       for _ in range(self.num_layers))

    def forward(self, X_0):
        if self.num_layers == 0:
            return X_0
        w = torch.cat([layer.weight.weight for layer in self.cross_net], dim=0)     # (L, d)
        b = torch.stack([layer.bias for layer in self.cross_net], dim=0)             # (L, d)
        return F2.crossnet(X_0, w, b)


class CrossNetV2(nn.Module):
    def __init__(self, input_dim, num_layers):
        super(CrossNetV2, self).__init__()
        self.num_layers = num_layers
        self.cross_layers = nn.ModuleList(nn.Linear(input_dim, input_dim)
                                          for _ in range(self.num_layers))

    def forward(self, X_0):
        X_i = X_0  # b x dim
        for i in range(self.num_layers):
            layer = self.cross_layers[i]
            lin = F2.linear_act(X_i, layer.weight, layer.bias, B2_ACT_NONE)
            X_i = X_i + X_0 * lin
        return X_i


# --------------------------------------------------------------------------------------
# CIN
# --------------------------------------------------------------------------------------
class CompressedInteractionNet(nn.Module):
    def __init__(self, num_fields, cin_hidden_units, output_dim=1):
        super(CompressedInteractionNet, self).__init__()
        self.cin_hidden_units = cin_hidden_units
        self.fc = nn.Linear(sum(cin_hidden_units), output_dim)
        self.cin_layer = nn.ModuleDict()
        for i, unit in enumerate(self.cin_hidden_units):
            in_channels = num_fields * self.cin_hidden_units[i - 1] if i > 0 else num_fields ** 2
            out_channels = unit
            self.cin_layer["layer_" + str(i + 1)] = nn.Conv1d(in_channels, out_channels, kernel_size=1)

    def forward(self, feature_emb):
        return F2.cin_forward(feature_emb,
                              [self.cin_layer["layer_" + str(i + 1)] for i in range(len(self.cin_hidden_units))],
                              self.fc)


# --------------------------------------------------------------------------------------
# Dice / MLP / DIN attention
# --------------------------------------------------------------------------------------
class Dice(nn.Module):
    def __init__(self, input_dim, eps=1e-9):
        super(Dice, self).__init__()
        self.bn = nn.BatchNorm1d(input_dim, affine=False, eps=eps, momentum=0.01)
        self.alpha = nn.Parameter(torch.zeros(input_dim))

    def forward(self, X):
        return F2.dice_forward(X, self.bn, self.alpha, self.training)


class MLP_Block(nn.Module):
    def __init__(self,
                 input_dim,
                 hidden_units=[],
                 hidden_activations="ReLU",
                 output_dim=None,
                 output_activation=None,
                 dropout_rates=0.0,
                 batch_norm=False,
                 bn_only_once=False,  # Set True for inference speed up
                 use_bias=True):
        super(MLP_Block, self).__init__()
        dense_layers = []
        if not isinstance(dropout_rates, list):
            dropout_rates = [dropout_rates] * len(hidden_units)
        if not isinstance(hidden_activations, list):
            hidden_activations = [hidden_activations] * len(hidden_units)
        hidden_activations = get_activation(hidden_activations, hidden_units)
        hidden_units = [input_dim] + hidden_units
        if batch_norm and bn_only_once:
            dense_layers.append(nn.BatchNorm1d(input_dim))
        for idx in range(len(hidden_units) - 1):
            dense_layers.append(nn.Linear(hidden_units[idx], hidden_units[idx + 1], bias=use_bias))
            if batch_norm and not bn_only_once:
                dense_layers.append(nn.BatchNorm1d(hidden_units[idx + 1]))
            if hidden_activations[idx]:
                dense_layers.append(hidden_activations[idx])
            if dropout_rates[idx] > 0:
                dense_layers.append(nn.Dropout(p=dropout_rates[idx]))
        if output_dim is not None:
            dense_layers.append(nn.Linear(hidden_units[-1], output_dim, bias=use_bias))
        if output_activation is not None:
            dense_layers.append(get_activation(output_activation))
        self.mlp = nn.Sequential(*dense_layers)  # * used to unpack list

    def forward(self, inputs):
        mods = list(self.mlp)
        x = inputs
        if F2.mlp_chain_supported() and inputs.is_cuda:
            # a pure Linear(+ReLU/Sigmoid) stack runs as ONE autograd node (cross-layer epilogue fusion)
            layers, i, pure = [], 0, len(mods) > 0
            while i < len(mods):
                if type(mods[i]) != nn.Linear:
                    pure = False
                    break
                act = B2_ACT_NONE
                if i + 1 < len(mods) and type(mods[i + 1]) == nn.ReLU:
                    act = B2_ACT_RELU
                elif i + 1 < len(mods) and type(mods[i + 1]) == nn.Sigmoid:
                    act = B2_ACT_SIGMOID
                layers.append((mods[i].weight, mods[i].bias, act))
                i += 2 if act != B2_ACT_NONE else 1
            if pure:
                return F2.mlp_chain(x, layers)
        i = 0
        while i < len(mods):
            m = mods[i]
            if type(m) == nn.Linear:
                act = B2_ACT_NONE
                if i + 1 < len(mods) and type(mods[i + 1]) == nn.ReLU:
                    act = B2_ACT_RELU
                elif i + 1 < len(mods) and type(mods[i + 1]) == nn.Sigmoid:
                    act = B2_ACT_SIGMOID
                x = F2.linear_act(x, m.weight, m.bias, act)   # Linear + activation fused
                i += 2 if act != B2_ACT_NONE else 1
            else:
                # BatchNorm1d / Dropout / PReLU / Tanh ...: applied as the reference does
                # (mlp_block.py:72-80); Dice dispatches to its own kernel.
                x = m(x)
                i += 1
        return x


class DIN_Attention(nn.Module):
    def __init__(self,
                 embedding_dim=64,
                 attention_units=[32],
                 hidden_activations="ReLU",
                 output_activation=None,
                 dropout_rate=0,
                 batch_norm=False,
                 use_softmax=False):
        super(DIN_Attention, self).__init__()
        self.embedding_dim = embedding_dim
        self.use_softmax = use_softmax
        if isinstance(hidden_activations, str) and hidden_activations.lower() == "dice":
            hidden_activations = [Dice(units) for units in attention_units]
        self.attention_layer = MLP_Block(input_dim=4 * embedding_dim,
                                         output_dim=1,
                                         hidden_units=attention_units,
                                         hidden_activations=hidden_activations,
                                         output_activation=output_activation,
                                         dropout_rates=dropout_rate,
                                         batch_norm=batch_norm)

    def forward(self, target_item, history_sequence, mask=None):
        return F2.din_attention(self, target_item, history_sequence, mask)
