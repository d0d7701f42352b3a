"""Multi-rank GPU parity of the data-parallel fine-tuning step ("grad equality across ranks and vs a single
big-batch run"): world = 2 `VaultTrainStep` -- segmented backward graphs, then per gradient range either NCCL all-reduce (bf16 and fp32
payload) + AdamW, or the fused NVSwitch-multicast kernel (in-switch fp32 reduce of the rank's slice + AdamW + multicast store of the new
weights, `vault_mc_adamw_step`) -- against ONE rank stepping the concatenated global batch.

Checked after 3 optimizer steps (dropout off, constant lr -- 1e-3 tiny / 2e-5 base -- so the first step already moves the weights):
  * the two ranks hold bit-identical weights (same reduced gradients, same update);
  * the mean of the ranks' losses equals the single-rank global-batch loss (CE mean over the global batch) within 5e-3;
  * the weight delta of the data-parallel run has cosine >= 0.999 with the single-rank delta (tiny model, fp32 payload: >= 0.9999).

Needs two visible B200s: skipped on a one-GPU box (run with `gpurun --gpus 2 -- python -m pytest tests/test_multigpu_gpu.py -m gpu`)."""
import json
import os
import socket
import tempfile

import pytest
import torch

pytestmark = pytest.mark.gpu

FWD = ("input_ids", "attention_mask", "token_type_ids", "pixel_values", "pixel_mask")
STEPS = 3
B_LOCAL = 4


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _build(dims_name, dev):
    from oracle import synth
    from oracle.ref_loader import hf_configs
    from vault_b200 import VaultForTMSC

    d = getattr(synth.Dims, dims_name)()
    sd = synth.make_state_dict(d, seed=0)
    vc, lc = hf_configs(d)
    m = VaultForTMSC(vc, n_classes=d.n_classes, vilt_dropout_prob=d.head_dropout, bert_config=lc)
    m.embeddings.text_embeddings.position_embedding_type = "NOT_absolute"
    m.load_state_dict(sd, strict=False)
    return d, m.to(dev).train()


def _flat_weights(m):
    eng = m.engine
    return eng.master[:eng.n_train].detach().clone()


def _worker(rank, world, port, dims_name, comm, comm_dtype, text_len, image_hw, lr, out_dir):
    import torch.distributed as dist

    from oracle import synth
    from vault_b200 import VaultTrainStep

    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=dev)
    try:
        solo = [dist.new_group([r]) for r in range(world)]  # one-rank groups (created collectively): a world-1 step inside this process
        d, m = _build(dims_name, dev)
        glob = [synth.make_inputs(d, batch=world * B_LOCAL, text_len=text_len, image_hw=image_hw, seed=11 + s, var_text=True) for s in range(STEPS)]
        keys = FWD + ("labels",)
        w0 = None
        # ---- data-parallel run: this rank's shard of every global batch ----
        try:
            ts = VaultTrainStep(m, lr=lr, total_steps=None, dropout=False, grad_comm_dtype=comm_dtype, comm=comm)
        except RuntimeError as e:
            if comm == "multimem" and "multicast support" in str(e):
                with open(os.path.join(out_dir, f"rank{rank}.json"), "w") as f:
                    json.dump(dict(skipped=str(e)), f)
                return
            raise
        assert ts.world == world and ts.overlap and (ts.mc is not None) == (comm == "multimem")
        w0 = _flat_weights(m)
        losses = []
        for g in glob:
            shard = {k: g[k][rank * B_LOCAL:(rank + 1) * B_LOCAL].contiguous() for k in keys}
            losses.append(ts.step(shard).loss())
        ts.synchronize()
        w_dp = _flat_weights(m)
        gathered = [torch.empty_like(w_dp) for _ in range(world)]
        dist.all_gather(gathered, w_dp)
        same = all(torch.equal(gathered[0], x) for x in gathered[1:])
        lt = torch.tensor(losses, device=dev, dtype=torch.float64)
        dist.all_reduce(lt)
        dp_loss = (lt / world).tolist()
        res = dict(rank=rank, identical_across_ranks=bool(same), dp_loss=dp_loss)
        if rank == 0:
            # ---- single-rank run on the concatenated batch, same initial weights ----
            d2, m2 = _build(dims_name, dev)
            ts1 = VaultTrainStep(m2, lr=lr, total_steps=None, dropout=False, process_group=solo[0])
            assert ts1.world == 1
            assert torch.equal(_flat_weights(m2), w0)
            l1 = [ts1.step({k: g[k] for k in keys}).loss() for g in glob]
            ts1.synchronize()
            w_1 = _flat_weights(m2)
            a, b = (w_dp - w0).double(), (w_1 - w0).double()
            res.update(single_loss=l1, delta_cosine=float(a @ b / (a.norm() * b.norm())), delta_rel=float((a - b).norm() / b.norm()),
                       delta_norm=float(b.norm()))
        with open(os.path.join(out_dir, f"rank{rank}.json"), "w") as f:
            json.dump(res, f)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("comm,comm_dtype", [("nccl", "bf16"), ("nccl", "fp32"), ("multimem", "fp32"), ("multimem", "bf16")])
@pytest.mark.parametrize("dims_name,text_len,image_hw,lr", [("tiny", 16, (64, 96), 1e-3), ("base", 40, (384, 384), 2e-5)])
def test_two_rank_step_matches_single_rank_global_batch(comm, comm_dtype, dims_name, text_len, image_hw, lr):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp

    with tempfile.TemporaryDirectory() as out:
        mp.spawn(_worker, args=(2, _free_port(), dims_name, comm, comm_dtype, text_len, image_hw, lr, out), nprocs=2, join=True)
        r0 = json.load(open(os.path.join(out, "rank0.json")))
        r1 = json.load(open(os.path.join(out, "rank1.json")))
    if "skipped" in r0:
        pytest.skip(r0["skipped"])
    print(json.dumps(dict(comm=comm, payload=comm_dtype, **{k: r0[k] for k in ("dp_loss", "single_loss", "delta_cosine", "delta_rel", "identical_across_ranks")})))
    assert r0["identical_across_ranks"] and r1["identical_across_ranks"]
    for a, b in zip(r0["dp_loss"], r0["single_loss"]):
        assert abs(a - b) <= 5e-3, (r0["dp_loss"], r0["single_loss"])  # the loss tolerance of tests/test_parity_gpu.py
    # measured on 2 x B200 (3 steps): tiny 0.99993 (bf16 payload) / 0.999997 (fp32); base 0.99929 / 0.99923 -- at full size the payload
    # dtype is below the bf16 compute noise (HF-AdamW without bias correction takes near-sign steps at first, so elements with a tiny
    # gradient flip with any rounding difference between the B=4 shards and the B=8 batch)
    assert r0["delta_cosine"] >= (0.9999 if (comm_dtype == "fp32" and dims_name == "tiny") else 0.999), r0
