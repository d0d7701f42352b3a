"""CPU tests of the host side: drop-in class surface, HF-compatible keys, flat-buffer layout, schedule, and the world_size-2
data-parallel reduction over gloo."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import synth, vault_oracle as O
from oracle.ref_loader import hf_configs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def tiny_model(**dkw):
    from vault_b200 import VaultForTMSC

    d = synth.Dims.tiny(**dkw)
    vc, lc = hf_configs(d)
    m = VaultForTMSC(vc, n_classes=d.n_classes, vilt_dropout_prob=0.1, bert_config=lc)
    return d, m


def test_state_dict_keys_are_hf_compatible():
    d, m = tiny_model()
    mine = {k: tuple(v.shape) for k, v in m.state_dict().items() if not k.endswith("position_ids") and not k.endswith("token_type_ids")}
    assert mine == synth.param_shapes(d), set(mine) ^ set(synth.param_shapes(d))


def test_reference_class_surface():
    import vault_b200
    from vault_b200.models.vault import VaultForMaskedLM, VaultForTMSC, VaultModel, VaultProcessor  # noqa: F401  reference import path

    d, m = tiny_model()
    assert isinstance(m, vault_b200.VaultModel) and m.bert is not None and m.freeze_lm is False
    assert m.embeddings.text_embeddings.position_embedding_type == "NOT_absolute"  # ref:vault/models/vault/model.py:77-79
    assert m.get_input_embeddings() is m.bert.get_input_embeddings()
    # the reference's typo'd attributes leave ViLT's real dropout at the config value (0.0)
    assert m.config.hidden_dropout_prob == 0.0 and m.config.t_prob == 0.1
    m.resize_token_embeddings(600)
    assert m.bert.get_input_embeddings().weight.shape[0] == 600
    from vault_b200.model import VaultMixin
    assert issubclass(VaultForMaskedLM, VaultMixin) and issubclass(VaultModel, VaultMixin)  # every class is a VaultMixin, as in the reference


def test_frozen_lm_and_no_lm_variants():
    from vault_b200 import VaultForTMSC, VaultModel

    d = synth.Dims.tiny()
    vc, lc = hf_configs(d)
    m = VaultModel(vc, bert_config=lc, freeze_lm=True)
    assert all(not p.requires_grad for p in m.bert.parameters()) and m.pooler is not None
    d0 = synth.Dims.tiny(lm_layers=0)
    vc0, _ = hf_configs(d0)
    m0 = VaultForTMSC(vc0, n_classes=3)
    assert m0.bert is None and m0.embeddings.text_embeddings.position_embedding_type == "absolute"


def test_cpu_call_fails_loudly():
    d, m = tiny_model()
    inp = synth.make_inputs(d, batch=1, text_len=8)
    with pytest.raises(RuntimeError, match="no CPU"):
        m(input_ids=inp["input_ids"], attention_mask=inp["attention_mask"], token_type_ids=inp["token_type_ids"], pixel_values=inp["pixel_values"])


def test_flat_layout_order_and_never_grad():
    d, m = tiny_model()
    eng = m.engine
    train, static = eng._param_order()
    assert set(static) == O.grads_never_set(d) == eng.never_grad_names()
    # reverse-topological: head first, LM embeddings last; q,k,v adjacent so the fused [3H,H] slice is contiguous
    assert train[0] == "classifier.1.weight" and train[-1].startswith("bert.embeddings.")
    for pre in ("encoder.layer.1.attention.attention.", "bert.encoder.layer.0.attention.self."):
        i = train.index(pre + "query.weight")
        assert train[i:i + 3] == [pre + n for n in ("query.weight", "key.weight", "value.weight")]
        j = train.index(pre + "query.bias")
        assert train[j:j + 3] == [pre + n for n in ("query.bias", "key.bias", "value.bias")]
    m.freeze_lm = True
    for p in m.bert.parameters():
        p.requires_grad_(False)
    train2, static2 = eng._param_order()
    assert all(not n.startswith("bert.") for n in train2) and sum(n.startswith("bert.") for n in static2) > 0


def test_schedule_matches_oracle():
    from vault_b200.train import VaultTrainStep

    ts = VaultTrainStep.__new__(VaultTrainStep)
    ts.lr, ts.total_steps, ts.warmup_ratio = 2e-5, 1000, 0.1
    for s in (0, 1, 50, 99, 100, 101, 555, 999, 1000):
        assert abs(ts.lr_at(s) - O.linear_warmup_lr(s, 1000, 2e-5)) < 1e-15
    assert ts.lr_at(0) == 0.0


def test_multicast_slices_partition_every_gradient_range():
    """Data-parallel multicast path (train.py): rank r owns slice r of every finished gradient range -- the slices are disjoint, cover the
    range, and start / end on multiples of 8 parameters (one 32-byte unit of vault_mc_adamw_step) wherever the range does."""
    from vault_b200.train import VaultTrainStep

    for world in (2, 3, 4, 8):
        for lo, hi in ((0, 64), (128, 128 + 8 * 1000), (64, 64 + 27_981_888), (0, 8), (192, 192 + 24)):
            cover = []
            for rank in range(world):
                ts = VaultTrainStep.__new__(VaultTrainStep)
                ts.world, ts.rank = world, rank
                a, b = ts._mc_slice(lo, hi)
                assert lo <= a <= b <= hi and (a - lo) % 8 == 0 and ((b - lo) % 8 == 0 or b == hi)
                cover.append((a, b))
            assert cover[0][0] == lo and cover[-1][1] == hi
            assert all(cover[i][1] == cover[i + 1][0] for i in range(world - 1))


def test_weight_gradient_tile_choice_fills_the_sms():
    """_wgrad_cfg: 128x256 tiles with a power-of-two split-K where that fills the 148 SMs, 128x192 for the 2304 x 768 QKV gradient
    (54 tiles x 2 = 108 CTAs -> 72 x 2 = 144)."""
    d, m = tiny_model()
    eng = m.engine
    eng.sms = 148
    assert eng._wgrad_cfg(2304, 768, 11808) == (192, 2)     # QKV
    assert eng._wgrad_cfg(3072, 768, 11808) == (256, 2)     # MLP-1: 72 tiles x 2
    assert eng._wgrad_cfg(768, 3072, 11808) == (256, 2)     # MLP-2
    assert eng._wgrad_cfg(768, 768, 11808) == (256, 8)      # attention output: 18 tiles x 8
    bn, split = eng._wgrad_cfg(768, 768, 100)               # two k-blocks only: no split
    assert split == 1 and bn in (128, 192, 256)
    # grouped launch of a layer's four weight gradients: 54 + 18 + 72 + 72 = 216 tiles -> x2 = 432 = 2.92 waves of 148 SMs
    assert eng._group_split(216, 185) == 2 and eng._group_split(216, 64) == 2
    assert eng._group_split(216, 3) == 1        # too few k-blocks to split
    assert eng._group_split(148, 100) == 1      # one full wave as it is
    assert eng._group_split(30, 100) == 4       # 120 of 148 SMs (0.81); eight splits would leave a 62 % second wave


def test_shadow_only_bitmap_marks_dense_matrices_only():
    """The multicast optimizer step leaves the fp32 masters of a 64-parameter block sharded iff its bit is set: set for the dense projection
    matrices of the encoder layers (read through the bf16 shadow only), clear for everything a kernel reads in fp32."""
    import types
    from vault_b200.engine import ALIGN, Slot, VaultEngine

    names = ["classifier.1.weight", "encoder.layer.1.attention.attention.query.bias", "encoder.layer.1.layernorm_before.weight",
             "encoder.layer.1.attention.output.dense.weight", "encoder.layer.1.attention.attention.query.weight",
             "encoder.layer.1.intermediate.dense.weight", "encoder.layer.1.output.dense.bias", "embeddings.position_embeddings",
             "bert.encoder.layer.0.attention.self.value.weight", "bert.encoder.layer.0.output.LayerNorm.weight",
             "bert.embeddings.word_embeddings.weight", "pooler.dense.weight"]
    off, slots = 0, {}
    for i, n in enumerate(names):
        numel = 64 * (i + 1) + (8 if i % 2 else 0)
        slots[n] = Slot(n, off, numel, (numel,), True)
        off += (numel + ALIGN - 1) // ALIGN * ALIGN
    fake = types.SimpleNamespace(slots=slots, n_train=off, device=torch.device("cpu"))
    words = VaultEngine.shadow_only_bitmap(fake)
    bits = [(int(words[b // 32]) >> (b % 32)) & 1 for b in range((off + ALIGN - 1) // ALIGN)]
    for n, sl in slots.items():
        dense = any(k in n for k in ("query.weight", "value.weight", "attention.output.dense.weight", "intermediate.dense.weight"))
        blocks = range(sl.off // ALIGN, (sl.off + sl.numel + ALIGN - 1) // ALIGN)
        assert all(bits[b] == int(dense) for b in blocks), n


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _dp_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from vault_b200.train import allreduce_flat_

    d = synth.Dims.tiny()
    sd = synth.make_state_dict(d, seed=0)
    full = synth.make_inputs(d, batch=4, text_len=12, seed=5, var_text=True)
    keys = ("input_ids", "attention_mask", "token_type_ids", "pixel_values", "pixel_mask")

    def grads(batch):
        params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        o = O.vault_forward(params, d, **{k: batch[k] for k in keys})
        O.ce_loss(O.tmsc_logits(params, d, o["pooler_output"]), batch["labels"]).backward()
        names = sorted(k for k, p in params.items() if p.grad is not None)
        return names, torch.cat([params[k].grad.flatten() for k in names])

    local = {k: v[rank * 2:(rank + 1) * 2] for k, v in full.items()}  # batch sharding: 2 rows per rank
    names, flat = grads(local)
    allreduce_flat_(flat, None, bucket_elems=100_000)  # several buckets
    flat *= 1.0 / world                                # what AdamW's grad_scale applies
    if rank == 0:
        _, ref = grads(full)
        torch.save(dict(err=(flat - ref).abs().max().item(), scale=ref.abs().max().item()), out)
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_gradient_average_world2_gloo(tmp_path):
    """Sharding the batch over 2 ranks + sum-all-reduce + 1/world == the single-process full-batch gradient (CE mean, equal shards)."""
    out = str(tmp_path / "dp.pt")
    mp.spawn(_dp_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r = torch.load(out)
    assert r["err"] <= 2e-5 * max(r["scale"], 1.0), r


def _tiny_cfgs():
    from transformers import BertConfig, ViltConfig

    vc = ViltConfig(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256, max_position_embeddings=40)
    lc = BertConfig(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256)
    return vc, lc


def test_head_wrappers_route_the_trunk_through_the_engine(monkeypatch):
    """The VaultFor* head classes (ref:vault/models/vault/model.py:375-509): HF head modules on top of `_trunk` (stubbed here -- no GPU),
    reference constructor extras (n_classes, num_images, widened modality table), HF-compatible keys, engine layout with the vilt. prefix."""
    from vault_b200.engine import VaultEngine
    from vault_b200.model import VaultMixin, _KernelTrunk
    from vault_b200.models.vault import (VaultForImageAndTextRetrieval, VaultForImagesAndTextClassification, VaultForMaskedLM,
                                         VaultForQuestionAnswering)

    calls = []

    def fake_trunk(self, input_ids=None, pixel_values=None, image_token_type_idx=None, **kw):
        B, T = input_ids.shape
        calls.append((type(self).__name__, image_token_type_idx, tuple(pixel_values.shape)))
        g = torch.Generator().manual_seed(len(calls))
        lhs = torch.randn(B, T + 5, 128, generator=g).requires_grad_(True)
        return lhs, torch.tanh(lhs[:, 0]), torch.ones(B, T + 5, dtype=torch.uint8)

    monkeypatch.setattr(VaultMixin, "_trunk", fake_trunk)
    from vault_b200.model import _KernelDecoder
    monkeypatch.setattr(_KernelDecoder, "forward", torch.nn.Linear.forward)  # the decoder GEMM is a kernel too: stubbed like the trunk (no GPU here)
    ids = torch.randint(5, 100, (2, 7))
    px = torch.randn(2, 3, 64, 64)
    common = dict(input_ids=ids, attention_mask=torch.ones_like(ids), token_type_ids=torch.zeros_like(ids))

    vc, lc = _tiny_cfgs()
    mlm = VaultForMaskedLM(vc, bert_config=lc)
    labels = torch.full((2, 7), -100)
    labels[0, 2] = 17
    out = mlm(**common, pixel_values=px, labels=labels)
    assert tuple(out.logits.shape) == (2, 7, vc.vocab_size) and torch.isfinite(out.loss)
    out.loss.backward()
    assert mlm.mlm_score.transform.dense.weight.grad is not None
    assert isinstance(mlm.vilt, _KernelTrunk) and "vilt.encoder.layer.0.attention.attention.query.weight" in mlm.state_dict()

    vc, lc = _tiny_cfgs()
    vqa = VaultForQuestionAnswering(vc, bert_config=lc, n_classes=11)
    out = vqa(**common, pixel_values=px, labels=torch.rand(2, 11))
    assert tuple(out.logits.shape) == (2, 11) and vqa.classifier[-1].out_features == 11 and torch.isfinite(out.loss)

    vc, lc = _tiny_cfgs()
    ret = VaultForImageAndTextRetrieval(vc, bert_config=lc)
    assert tuple(ret(**common, pixel_values=px).logits.shape) == (2, 1)

    vc, lc = _tiny_cfgs()
    nlvr = VaultForImagesAndTextClassification(vc, bert_config=lc)
    assert nlvr.config.num_images == 2 and nlvr.vilt.embeddings.token_type_embeddings.weight.shape[0] == 3
    calls.clear()
    out = nlvr(**common, pixel_values=torch.randn(2, 2, 3, 64, 64), labels=torch.tensor([0, 1]))
    assert [c[1] for c in calls] == [1, 2] and calls[0][2] == (2, 3, 64, 64) and tuple(out.logits.shape) == (2, 2)

    # engine layout for a wrapper: trunk names carry "vilt.", head parameters stay outside the kernel-owned (trainable) range
    eng = VaultEngine(nlvr)
    train, static = eng._param_order()
    assert eng.vp == "vilt." and eng._k("layernorm.weight") == "vilt.layernorm.weight" and eng._k("bert.embeddings.LayerNorm.bias").startswith("bert.")
    assert train[0] == "vilt.pooler.dense.weight" and all(n.startswith(("vilt.", "bert.")) for n in train)
    assert {"classifier.0.weight", "classifier.3.bias", "vilt.embeddings.text_embeddings.word_embeddings.weight"} <= set(static)


def test_bench_workloads_are_consistent_with_baseline_shapes():
    """bench.py's named workloads: ViLT sequence = text + CLS + patches, patch grid = image / 32, FLOPs per sample = the closed form of
    SURVEY.md section 8d (3x forward, patch projection 2x, frozen LM 1x), and the oracle dims of each workload exist."""
    import bench

    assert bench.parse.__module__ == "bench" and set(bench.WORKLOADS) == {"config3", "target", "config4", "config5"}
    for name, w in bench.WORKLOADS.items():
        T, (hi, wi) = w["text_len"], w["image"]
        P = (hi // 32) * (wi // 32)
        assert w["patches"] == P and w["seq_len"] == T + 1 + P
        enc = lambda n: 12 * (14155776 * n + 3072 * n * n)  # 12 layers: 24 n H^2 + 4 n^2 H
        lm, vilt, patch = enc(T), enc(T + 1 + P), 4718592 * P
        want = ((1 if w["freeze_lm"] else 3) * lm + 3 * vilt + 2 * patch) / 1e9
        assert abs(w["train_gflop"] - want) / want < 2e-3, (name, w["train_gflop"], want)
        assert abs(bench.train_gflop_valid_tokens([T] * 3, P, w["freeze_lm"]) - want) / want < 1e-9  # all-valid text = the dense-shape figure
        assert bench.train_gflop_valid_tokens([T // 2] * 3, P, w["freeze_lm"]) < want
        cfg = bench.workload_config(name)
        assert cfg["per_gpu_batch"] == 32 and cfg["text_len"] == T
        lc = bench._hf_lm_config(name)
        assert (lc.model_type == "roberta") == (w["lm_kind"] == "roberta") and (lc.vocab_size == 64001) == (w["lm_kind"] == "roberta")
        # one `config` for both arms of the bench (the driver compares them): workload-only keys, nothing about how an arm runs it
        c1 = bench.bench_config(name, 32, 2)
        assert c1["global_batch"] == 64 and c1["parallelism"] == "dp2" and c1["name"] == name and "cuda_graph" not in c1 and "model" not in c1


def test_mlm_decoder_class_swap_keeps_the_state_dict():
    from transformers import BertConfig, ViltConfig

    from vault_b200.model import _KernelDecoder
    from vault_b200.models.vault import VaultForMaskedLM

    kw = dict(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=512, vocab_size=510)
    m = VaultForMaskedLM(ViltConfig(**kw), bert_config=BertConfig(**kw))
    assert isinstance(m.mlm_score.decoder, _KernelDecoder)
    keys = set(m.state_dict())
    assert {"mlm_score.decoder.weight", "mlm_score.decoder.bias", "mlm_score.transform.dense.weight"} <= keys
    with pytest.raises(RuntimeError, match="CUDA"):  # no CPU path, as everywhere else
        m.mlm_score.decoder(torch.zeros(1, 2, 128))


def _check_bench_line(d, reference=False):
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "e2e"):
        assert k in d, k
    assert d["metric"] == "VaultModel fine-tune samples/sec" and d["unit"] == "samples/s" and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
    cb = d["cpu_baseline"]
    assert cb is None or {"value", "unit", "cores", "kind", "sample"} <= set(cb)
    if reference:
        assert d["impl"] == "reference" and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"] and cb["kind"] in ("port", "reference")
    else:
        r = d["roofline"]
        assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] in ("hbm", "tensor")
        assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and (r["traffic"] is None or isinstance(r["traffic"], (int, float)))
        assert d["gpu_launches"] > 0 and {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"]) and d["e2e"]["h2d_bytes_per_step"] > 0


def test_committed_bench_lines_follow_the_contract():
    """The measured JSON lines kept under profiles/ (what bench.py printed on the B200) carry every key of the bench contract."""
    import glob
    import json

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = sorted(glob.glob(os.path.join(root, "profiles", "r0[12]_bench_*gpu*.json")))
    assert files
    for f in files:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        _check_bench_line(d, reference=d.get("impl") == "reference")


def test_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the reference's own VaultForTMSC on the host cores; oracle port only where the reference files are absent)
    end to end: one JSON line, same metric / unit, and a `config` identical to the one this repo's arm prints for the same command."""
    import json
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--batch", "2"], capture_output=True,
                       text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    _check_bench_line(d, reference=True)
    import bench
    from oracle import ref_loader

    assert d["config"] == bench.bench_config("target", 2, 1) and d["value"] > 0  # default workload = the north-star target shape
    assert d["cpu_baseline"]["kind"] == ("reference" if ref_loader.available() else "port")
    # under torchrun only rank 0 runs it; the other ranks exit 0 without work
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], capture_output=True,
                       text=True, timeout=120, cwd=root, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_processor_falls_back_to_the_stock_vilt_processor_and_swaps_the_tokenizer(monkeypatch):
    """VaultProcessor.from_pretrained (ref:vault/models/vault/processor.py:7-18): checkpoints without processor files load the stock ViLT-B/32
    processor, the LM's tokenizer replaces ViLT's; with nothing loadable the first error surfaces."""
    import transformers

    from vault_b200 import processor as P

    calls = []

    class FakeProc:
        tokenizer = "vilt-tokenizer"

    def fake_from_pretrained(cls, src, *a, **k):
        calls.append(src)
        if src != "dandelin/vilt-b32-mlm":
            raise OSError(f"no processor files in {src}")
        return FakeProc()

    monkeypatch.setattr(transformers.ViltProcessor, "from_pretrained", classmethod(fake_from_pretrained))
    monkeypatch.setattr(P.AutoTokenizer, "from_pretrained", staticmethod(lambda src: "tok:" + src))
    p = P.VaultProcessor.from_pretrained("some/vilt-checkpoint", "vinai/bertweet-base")
    assert calls == ["some/vilt-checkpoint", "dandelin/vilt-b32-mlm"] and p.tokenizer == "tok:vinai/bertweet-base"
    assert P.VaultProcessor.from_pretrained("dandelin/vilt-b32-mlm").tokenizer == "vilt-tokenizer"
    monkeypatch.setattr(P, "_FALLBACK_PROCESSORS", ())
    with pytest.raises(OSError, match="some/other"):
        P.VaultProcessor.from_pretrained("some/other")


def test_from_pretrained_with_local_checkpoints(tmp_path):
    """VaultMixin.from_pretrained (ref:vault/models/vault/model.py:92-128) on local ViLT / BERT checkpoints written by save_pretrained: trunk and LM
    weights arrive under the HF keys, the head is freshly initialised, freeze_lm / the position-embedding gate / n_classes behave as in the reference."""
    from transformers import BertConfig, BertModel, ViltConfig, ViltModel

    from vault_b200.models.vault import VaultForTMSC, VaultModel

    kw = dict(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=512)
    vd, bd = str(tmp_path / "vilt"), str(tmp_path / "bert")
    torch.manual_seed(0)
    v = ViltModel(ViltConfig(vocab_size=512, **kw))
    v.save_pretrained(vd)
    b = BertModel(BertConfig(vocab_size=512, max_position_embeddings=64, **kw), add_pooling_layer=False)
    b.save_pretrained(bd)
    m = VaultForTMSC.from_pretrained(vd, bd, freeze_lm=True, n_classes=3, vilt_dropout_prob=0.1)
    assert m.freeze_lm and m.classifier[1].out_features == 3 and m.embeddings.text_embeddings.position_embedding_type == "NOT_absolute"
    sd = m.state_dict()
    assert torch.equal(sd["encoder.layer.1.output.dense.weight"], v.state_dict()["encoder.layer.1.output.dense.weight"])
    assert torch.equal(sd["bert.encoder.layer.0.attention.self.query.weight"], b.state_dict()["encoder.layer.0.attention.self.query.weight"])
    assert all(not p.requires_grad for p in m.bert.parameters()) and all(p.requires_grad for p in m.encoder.parameters())
    assert not any(k.startswith("bert.pooler") for k in sd)  # add_pooling_layer=False
    plain = VaultModel.from_pretrained(vd)
    assert plain.bert is None and plain.embeddings.text_embeddings.position_embedding_type == "absolute"
    keep_pos = VaultModel.from_pretrained(vd, bd, use_vilt_position_embeddings=True)
    assert keep_pos.bert is not None and not keep_pos.freeze_lm and keep_pos.embeddings.text_embeddings.position_embedding_type == "absolute"
    # the engine is built lazily from whatever LM is attached now
    assert keep_pos.engine.lm is keep_pos.bert and plain.engine.lm is None
