"""vault_b200 -- B200-native (sm_100a) implementation of VAuLT's data-parallel hot path behind the reference's class API."""
__all__ = ["VaultModel", "VaultForTMSC", "VaultProcessor", "VaultTrainStep"]


def __getattr__(name):  # lazy: importing the package must not import transformers / torch eagerly
    if name in ("VaultModel", "VaultForTMSC"):
        from . import model

        return getattr(model, name)
    if name == "VaultProcessor":
        from .processor import VaultProcessor

        return VaultProcessor
    if name == "VaultTrainStep":
        from .train import VaultTrainStep

        return VaultTrainStep
    raise AttributeError(name)
