"""VaultProcessor: the ViLT processor carrying the language model's tokenizer (ref:vault/models/vault/processor.py:7-18).

Host-side pre-processing at the boundary of the hot path -- the API is the reference's (`VaultProcessor.from_pretrained(vilt_directory,
bert_directory=None)` returns a `ViltProcessor` whose tokenizer is the LM's), the arithmetic on pixels can be moved to the GPU with
`gpu_images()` (vault_b200.image_processing.ViltImageProcessorB200, bit-identical to the CPU pipeline)."""
from typing import Optional, Sequence

from transformers import AutoTokenizer, ViltProcessor

# Checkpoints without processor files fall back to the stock ViLT-B/32 processor, which is what the reference does (ref :11-15).
_FALLBACK_PROCESSORS: Sequence[str] = ("dandelin/vilt-b32-mlm",)


def _load_vilt_processor(base_cls, source: str):
    """First processor that loads among `source` and the fallbacks; the error of `source` is re-raised if none does."""
    first_error = None
    for candidate in (source, *_FALLBACK_PROCESSORS):
        try:
            return base_cls.from_pretrained(candidate)
        except Exception as err:  # missing files, offline hub, malformed config: try the next candidate
            first_error = first_error or err
    raise first_error


class VaultProcessor(ViltProcessor):
    """`ViltProcessor` for VAuLT models: images through ViLT's image processor, text through the LM's tokenizer when one is attached."""

    @classmethod
    def from_pretrained(cls, vilt_directory: str, bert_directory: Optional[str] = None):
        processor = _load_vilt_processor(super(VaultProcessor, cls), vilt_directory)
        if bert_directory:
            processor.tokenizer = AutoTokenizer.from_pretrained(bert_directory)
        return processor

    @staticmethod
    def gpu_images(**kwargs):
        """The image half of the processor as CUDA kernels (resize, rescale, normalise, pad, pixel mask): same outputs, bit for bit."""
        from .image_processing import ViltImageProcessorB200

        return ViltImageProcessorB200(**kwargs)
