"""Test-harness stand-in for omegaconf (see hydra stub)."""


class DictConfig(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


class OmegaConf:
    @staticmethod
    def create(obj):
        if isinstance(obj, dict):
            return DictConfig({k: OmegaConf.create(v) for k, v in obj.items()})
        return obj
