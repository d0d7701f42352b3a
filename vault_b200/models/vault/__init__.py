"""Import path that mirrors the reference package layout: ``from vault_b200.models.vault import VaultModel, VaultForTMSC,
VaultProcessor`` (ref:vault/models/vault/__init__.py:1-22)."""
from ...model import (VaultForImageAndTextRetrieval, VaultForImagesAndTextClassification, VaultForMaskedLM,  # noqa: F401
                      VaultForQuestionAnswering, VaultForTMSC, VaultModel)
from ...processor import VaultProcessor  # noqa: F401
