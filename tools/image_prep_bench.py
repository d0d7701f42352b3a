"""Throughput of the GPU image pre-processing vs the CPU pipeline it replaces (Pillow bicubic + numpy normalise), one JSON line."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import image_oracle as IO
from vault_b200.image_processing import ViltImageProcessorB200
from vault_b200 import _abi

rng = np.random.default_rng(0)
B = 32
ims = [rng.integers(0, 256, (480, 640, 3), dtype=np.uint8) if i % 2 else rng.integers(0, 256, (500, 375, 3), dtype=np.uint8) for i in range(B)]
proc = ViltImageProcessorB200(device="cuda:0")
out = proc(ims); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10): out = proc(ims)
torch.cuda.synchronize()
e2e_ms = (time.perf_counter() - t0) / 10 * 1e3
# kernels only: replay the same launch with device-resident inputs
p = proc.plan([a.shape[:2] for a in ims]); dev = torch.device("cuda:0")
src = torch.from_numpy(np.concatenate([a.reshape(-1) for a in ims])).to(dev)
descs = torch.from_numpy(p["descs"].view(np.uint8).copy()).to(dev); coefs = torch.from_numpy(p["coefs"]).to(dev); bounds = torch.from_numpy(p["bounds"]).to(dev)
tmp = torch.empty(p["tmp_bytes"], dtype=torch.uint8, device=dev)
pv = torch.empty((B, 3, p["Hmax"], p["Wmax"]), device=dev); pm = torch.empty((B, p["Hmax"], p["Wmax"]), dtype=torch.int64, device=dev)
f = lambda: _abi.call("vault_image_preprocess", src.data_ptr(), descs.data_ptr(), coefs.data_ptr(), bounds.data_ptr(), tmp.data_ptr(), proc._lut.data_ptr(),
                      pv.data_ptr(), pm.data_ptr(), B, p["Hmax"], p["Wmax"], p["max_h_in"], p["max_w_out"], torch.cuda.current_stream().cuda_stream)
for _ in range(3): f()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(20): f()
e1.record(); torch.cuda.synchronize()
k_ms = e0.elapsed_time(e1) / 20
t0 = time.perf_counter(); IO.reference_preprocess(ims); cpu_ms = (time.perf_counter() - t0) * 1e3
bytes_alg = src.numel() + pv.numel() * 4 + pm.numel() * 8
print(json.dumps(dict(workload=f"{B} images 480x640 / 500x375 uint8 -> [B,3,{p['Hmax']},{p['Wmax']}] fp32 + mask", gpu_kernels_ms=round(k_ms, 4),
                      gpu_e2e_ms_incl_h2d=round(e2e_ms, 3), cpu_pillow_numpy_ms_1thread=round(cpu_ms, 1), images_per_s_gpu_e2e=round(B / e2e_ms * 1e3),
                      images_per_s_cpu=round(B / cpu_ms * 1e3), kernel_algorithmic_gbs=round(bytes_alg / k_ms / 1e6, 1))))
