"""VaultProcessor: ViltProcessor with the LM's tokenizer swapped in (ref:vault/models/vault/processor.py:7-18).  CPU-side
pre-processing at the boundary of the hot path: API kept, no kernels."""
from typing import Optional

from transformers import AutoTokenizer, ViltProcessor


class VaultProcessor(ViltProcessor):
    @classmethod
    def from_pretrained(cls, vilt_directory: str, bert_directory: Optional[str] = None):
        try:
            processor = super().from_pretrained(vilt_directory)
        except Exception:  # not all checkpoints ship a processor (ref :11-15)
            processor = super().from_pretrained("dandelin/vilt-b32-mlm")
        if bert_directory is not None:
            processor.tokenizer = AutoTokenizer.from_pretrained(bert_directory)
        return processor
