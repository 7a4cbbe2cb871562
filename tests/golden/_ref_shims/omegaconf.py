"""Stand-in for omegaconf (generator-only): attribute-access read-only mapping."""


class DictConfig(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


class OmegaConf:
    @staticmethod
    def create(d):
        return DictConfig(d)
