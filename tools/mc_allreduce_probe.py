"""Round-2 experiment driver (NOT on any product path; written at the end of round 1 without multi-GPU minutes left, so UNRUN so far):
the stand-alone NVSwitch-multicast all-reduce of tools/micro/multimem_allreduce.cu against NCCL on the gradient payload of the step.

    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -shared -Xcompiler -fPIC -o tools/micro/libmm_allreduce.so tools/micro/multimem_allreduce.cu
    timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/mc_allreduce_probe.py

Checks (1) the result against NCCL's on the same data, (2) time per all-reduce of 197 M bf16 gradients (394 MB) for several CTA counts, both
with CUDA events as the max over ranks.  Every cross-rank wait is the symmetric-memory handle's stream-ordered barrier with a timeout."""
import ctypes as C
import json
import os
import sys

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
lib = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "micro", "libmm_allreduce.so"))
lib.mm_allreduce_bf16.argtypes = [C.c_uint64, C.c_longlong, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_void_p]
lib.mm_allreduce_bf16.restype = C.c_int

n = int(sys.argv[1]) if len(sys.argv) > 1 else 197_017_344  # bf16 elements (multiple of 8)
buf = symm_mem.empty(n, dtype=torch.bfloat16, device=dev)
hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
mc = int(hdl.multicast_ptr)
if mc == 0:
    if rank == 0:
        print(json.dumps(dict(error="no multicast support on this system (multicast_ptr == 0)")))
    dist.destroy_process_group()
    sys.exit(0)
st = lambda: torch.cuda.current_stream().cuda_stream
BARRIER_MS = 20000


def mm_allreduce(ctas):
    hdl.barrier(channel=0, timeout_ms=BARRIER_MS)  # every rank's buffer is written
    rc = lib.mm_allreduce_bf16(mc, 0, n * 2, rank, world, ctas, st())
    assert rc == 0, rc
    hdl.barrier(channel=1, timeout_ms=BARRIER_MS)  # every slice is reduced and broadcast


# ---- (1) correctness against NCCL ----
g = torch.Generator(device=dev).manual_seed(100 + rank)
src = (torch.randn(n, device=dev, generator=g) * 0.1).to(torch.bfloat16)
ref = src.clone()
dist.all_reduce(ref)
buf.copy_(src)
mm_allreduce(16)
torch.cuda.synchronize()
err = (buf.float() - ref.float()).abs().max().item()
scale = ref.float().abs().max().item()

# ---- (2) timing, max over ranks ----
def timed(fn, reps=10):
    fn(); torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

res = dict(world=world, elems=n, mbytes=n * 2 / 1e6, max_abs_err=err, ref_max=scale, ok=bool(err <= 2 ** -7 * scale))
res["nccl_ms"] = timed(lambda: dist.all_reduce(ref))
for ctas in (4, 8, 16, 32, 64):
    res[f"multimem_{ctas}ctas_ms"] = timed(lambda: mm_allreduce(ctas))
if rank == 0:
    print(json.dumps(res), flush=True)
dist.destroy_process_group()
