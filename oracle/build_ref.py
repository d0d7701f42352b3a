"""Recipe for ``oracle/_ref/``: the reference's OWN source files for the hot path and its direct caller, taken from where they
lie under ``/root/reference`` so that the GPU box (which has no ``/root/reference``) can run the real reference on its host cores.

TEST / BENCH INFRASTRUCTURE.  ``oracle/_ref/`` is build output: git-ignored (never part of the history), NOT gpurun-ignored (it
travels to the GPU box like the built ``.so``).  Nothing under ``vault_b200/`` imports it.

    python -m oracle.build_ref          # also run by __graft_entry__.build() whenever /root/reference is present

Files (all of them pure Python; the arithmetic they drive lives in the installed ``transformers`` / ``torch``):

    vault/models/vault/model.py   -> oracle/_ref/vault_model.py    VaultModel / VaultForTMSC / the other VaultFor* heads
    vault/tmsc_utils/trainer.py   -> oracle/_ref/tmsc_trainer.py   Twitter201XTrainer (loss bookkeeping, metrics, schedule)
    vault/train_utils.py          -> oracle/_ref/train_utils.py    EarlyStopping (checkpoint save / load round trip)
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF_ROOT = os.environ.get("VAULT_REFERENCE_ROOT", "/root/reference")
FILES = {
    "vault/models/vault/model.py": "vault_model.py",
    "vault/tmsc_utils/trainer.py": "tmsc_trainer.py",
    "vault/train_utils.py": "train_utils.py",
}


def available() -> bool:
    return all(os.path.exists(os.path.join(REF_ROOT, src)) for src in FILES)


def build() -> dict:
    """Copy the files (byte-identical) and write a manifest with their SHA-256; returns the manifest."""
    if not available():
        raise RuntimeError(f"reference not present at {REF_ROOT}")
    os.makedirs(OUT, exist_ok=True)
    manifest = {"reference_root": REF_ROOT, "files": {}}
    for src, dst in FILES.items():
        s, d = os.path.join(REF_ROOT, src), os.path.join(OUT, dst)
        shutil.copyfile(s, d)
        with open(d, "rb") as f:
            manifest["files"][dst] = {"source": src, "sha256": hashlib.sha256(f.read()).hexdigest()}
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    return manifest


if __name__ == "__main__":
    print(json.dumps(build(), indent=1))
