"""Minimal CfgNode: attribute dict with merge_from_file / merge_from_list (test-only stub)."""
import yaml


class CfgNode(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def merge_from_file(self, path):
        with open(path) as f:
            self.update(yaml.safe_load(f) or {})

    def merge_from_list(self, lst):
        for k, v in zip(lst[0::2], lst[1::2]):
            self[k] = type(self[k])(v) if k in self and not isinstance(self[k], str) else v

    def clone(self):
        return CfgNode(self)
