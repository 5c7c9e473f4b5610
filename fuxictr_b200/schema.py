"""Feature schema that configures the fused gather (duck-type twin of the reference's
``FeatureMap``; the reference object itself is accepted everywhere one of these is).

Only the attributes the hot path reads exist here — ``features`` (ordered name -> spec),
``labels``, ``num_fields``, ``input_length``, ``default_emb_dim``, ``column_index`` — plus
the two aggregate queries model constructors call (``sum_emb_out_dim``,
``get_num_fields``; reference: fuxictr/features.py:113-154) and the batch-matrix column
layout of the collator (fuxictr/features.py:156-180,
fuxictr/pytorch/dataloaders/npz_dataloader.py:111-125).
"""
import json
from collections import OrderedDict

_WIDTH_KEY = {"sequence": "max_len", "embedding": "pretrain_dim"}


def _as_list(x):
    return x if isinstance(x, list) else [x]


class FeatureMap(object):
    def __init__(self, dataset_id="synthetic", data_dir=""):
        self.dataset_id, self.data_dir = dataset_id, data_dir
        self.features, self.labels, self.column_index = OrderedDict(), [], {}
        self.num_fields = self.total_features = self.input_length = 0
        self.group_id = self.default_emb_dim = None

    # ---- builders -----------------------------------------------------------------------
    @classmethod
    def from_specs(cls, specs, labels=("label",), embedding_dim=None, dataset_id="synthetic"):
        """specs: iterable of (feature name, spec dict) in batch-matrix column order."""
        self = cls(dataset_id)
        self.features = OrderedDict((name, dict(spec)) for name, spec in specs)
        self.labels = list(labels)
        self.default_emb_dim = embedding_dim
        self._finalise()
        return self

    def load(self, json_file, params):
        """Reads the reference's feature_map.json (a list of one-key dicts under "features")."""
        with open(json_file, "r", encoding="utf-8") as fd:
            blob = json.load(fd)
        if blob["dataset_id"] != self.dataset_id:
            raise RuntimeError("dataset_id={} does not match feature_map!".format(self.dataset_id))
        feats = OrderedDict()
        for entry in blob["features"]:
            feats.update(entry)
        self.num_fields = sum(1 for s in feats.values() if s["type"] != "meta")
        keep = params.get("use_features")
        self.features = OrderedDict((k, feats[k]) for k in keep) if keep else feats
        for override in params.get("feature_specs") or []:
            for name in _as_list(override["name"]):
                self.features[name].update({k: v for k, v in override.items() if k != "name"})
        self.labels = blob.get("labels", [])
        self.total_features = blob.get("total_features", 0)
        self.group_id = params.get("group_id")
        self.default_emb_dim = params.get("embedding_dim")
        self.set_column_index()

    def _finalise(self):
        self.num_fields = self.get_num_fields()
        self.total_features = sum(s.get("vocab_size", 0) for s in self.features.values())
        self.set_column_index()

    # ---- queries ------------------------------------------------------------------------
    def _selected(self, feature_source):
        wanted = _as_list(feature_source)
        for name, spec in self.features.items():
            if spec["type"] != "meta" and (not wanted or spec.get("source") in wanted):
                yield name, spec

    def get_num_fields(self, feature_source=[]):
        return sum(1 for _ in self._selected(feature_source))

    def sum_emb_out_dim(self, feature_source=[]):
        dflt = self.default_emb_dim
        return sum(spec.get("emb_output_dim", spec.get("embedding_dim", dflt))
                   for _, spec in self._selected(feature_source))

    def set_column_index(self):
        cursor = 0
        self.column_index = {}
        for name, spec in self.features.items():
            width_key = _WIDTH_KEY.get(spec["type"])
            if width_key is None:
                self.column_index[name] = cursor
                cursor += 1
            else:
                self.column_index[name] = list(range(cursor, cursor + spec[width_key]))
                cursor += spec[width_key]
        self.input_length = cursor
        for offset, label in enumerate(self.labels):
            self.column_index[label] = cursor + offset

    def get_column_index(self, feature):
        if feature not in self.column_index:
            self.set_column_index()
        return self.column_index[feature]

    def batch_views(self, batch_matrix):
        """Like batch_dict, but sequence features are column-RANGE views too (their columns are
        consecutive), so every entry aliases `batch_matrix`: what a static, graph-captured input
        buffer needs (pipeline.TrainPipeline).  The fused gather takes the row stride as is."""
        out = {}
        for c in list(self.features.keys()) + list(self.labels):
            idx = self.get_column_index(c)
            if isinstance(idx, list):
                if idx != list(range(idx[0], idx[0] + len(idx))):
                    raise ValueError("feature %s: columns %s are not consecutive" % (c, idx))
                out[c] = batch_matrix[:, idx[0]:idx[0] + len(idx)]
            else:
                out[c] = batch_matrix[:, idx]
        return out

    def batch_dict(self, batch_matrix):
        """name -> column view(s) of one (B, input_length + n_labels) matrix, as the collator
        hands them to the model (scalar features are strided views, sequences are copies)."""
        cols = list(self.features.keys()) + list(self.labels)
        return {c: batch_matrix[:, self.get_column_index(c)] for c in cols}
