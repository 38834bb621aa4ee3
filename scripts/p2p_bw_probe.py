"""One process, two GPUs of the box: device-to-device copy bandwidth and peer access, as evidence of the fabric the sharded step's
peer-memory stores and loads travel over (NVML's NVLink byte counters are not exposed in this container: NVML_ERROR_NOT_SUPPORTED)."""
import json
import sys

import torch

n = torch.cuda.device_count()
out = {"gpus": n}
if n >= 2:
    out["can_access_peer_0_1"] = bool(torch.cuda.can_device_access_peer(0, 1))
    x = torch.empty(1 << 28, dtype=torch.uint8, device="cuda:0")
    y = torch.empty(1 << 28, dtype=torch.uint8, device="cuda:1")
    for _ in range(3):
        y.copy_(x)
    torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.device(0):
        e0.record()
        for _ in range(10):
            y.copy_(x, non_blocking=True)
        e1.record()
        e1.synchronize()
    out["copy_0_to_1_GBps"] = 10 * x.numel() / (e0.elapsed_time(e1) * 1e-3) / 1e9
json.dump(out, sys.stdout)
print()
