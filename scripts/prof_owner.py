#!/usr/bin/env python
"""Per-CTA phase times of the one-barrier resident kernel (wr_debug_epoch_trace stamps of every CTA) on the bench shape."""
import json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from whisprrec_b200 import _lib
nU, nI, D, B, N = 6040, 3706, 64, 2048, 668862
dev = torch.device('cuda')
g = torch.Generator(device=dev); g.manual_seed(1)
P = torch.randn((nU + nI, D), device=dev, generator=g) * 0.1
M, V, G = torch.zeros_like(P), torch.zeros_like(P), torch.zeros_like(P)
ws = _lib.Workspace(dev)
ids = torch.stack([torch.randint(0, nU, (N,), device=dev, generator=g), torch.randint(0, nI, (N,), device=dev, generator=g),
                   torch.randint(1, nI, (N,), device=dev, generator=g)])
steps = (N + B - 1) // B
losses = torch.zeros(steps, device=dev)
lib = _lib.load()
k = [0]
def epoch():
    _lib.bprmf_epoch(P, M, V, G, ids, B, nU, k[0], 1e-3, 1e-6, losses, ws)
    k[0] += steps
epoch(); torch.cuda.synchronize()
ms = []
for _ in range(7):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); epoch(); e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
out = {'us_per_step': float(np.median(ms)) * 1e3 / steps}
trace = torch.zeros((steps, 8), dtype=torch.int64, device=dev)
ctr = torch.zeros((steps, 256, 4), dtype=torch.int64, device=dev)
lib.wr_debug_epoch_trace(trace.data_ptr(), ctr.data_ptr())
epoch(); torch.cuda.synchronize()
lib.wr_debug_epoch_trace(None, None)
sms = torch.cuda.get_device_properties(0).multi_processor_count
c = ctr.cpu().numpy().reshape(-1)[:steps * sms * 4].reshape(steps, sms, 4)[10:steps - 2].astype(np.int64)
fallback = (c[:, :, 0] & 1).mean()
c[:, :, 0] &= ~1
med = lambda x: float(np.median(x))
out.update({
    'fallback_scan_share': float(fallback),
    'process_median': med(c[:, :, 1] - c[:, :, 0]), 'process_max_cta': med((c[:, :, 1] - c[:, :, 0]).max(1)),
    'adam_median': med(c[:, :, 2] - c[:, :, 1]), 'adam_max_cta': med((c[:, :, 2] - c[:, :, 1]).max(1)),
    'barrier_median': med(c[:, :, 3] - c[:, :, 2]), 'barrier_min_cta': med((c[:, :, 3] - c[:, :, 2]).min(1)),
    'top_wait_median': med(c[1:, :, 0] - c[:-1, :, 3]), 'top_wait_max_cta': med((c[1:, :, 0] - c[:-1, :, 3]).max(1)),
    'arrival_skew': med(c[:, :, 2].max(1) - c[:, :, 2].min(1)), 'top_skew': med(c[:, :, 0].max(1) - c[:, :, 0].min(1)),
    'pass_skew': med(c[:, :, 3].max(1) - c[:, :, 3].min(1)),
    'last_arrival_to_first_pass': med(c[:, :, 3].min(1) - c[:, :, 2].max(1)),
    'last_arrival_to_last_pass': med(c[:, :, 3].max(1) - c[:, :, 2].max(1)),
    'step': med(np.diff(c[:, 0, 0]))})
print(json.dumps(out, indent=1))
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], 'w'), indent=1)
