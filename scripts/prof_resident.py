#!/usr/bin/env python
"""Where the time of the resident BPRMF kernel (csrc/epoch_kernel.cu) goes on the bench shape.

    python scripts/prof_resident.py [out.json]

1. one launch per epoch (327 steps, ids on the device): device time per step + the %globaltimer phase breakdown
   (wr_debug_epoch_trace) of CTA 0;
2. one launch per step: back to back (tables L2-resident) and with the L2 flushed between steps, beside the event floor;
3. host-fed streaming (wr_bprmf_ctx_*): latency of one step with the host waiting for completion (pinned in place and
   pageable ids), throughput with the host running ahead, and the GPU-side stamps of the same steps.
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from whisprrec_b200 import _lib  # noqa: E402

nU, nI, D, B, N = 6040, 3706, 64, 2048, 668862
dev = torch.device('cuda')
out = {}
g = torch.Generator(device=dev); g.manual_seed(1)
P = torch.randn((nU + nI, D), device=dev, generator=g) * 0.1
M, V, G = torch.zeros_like(P), torch.zeros_like(P), torch.zeros_like(P)
ws = _lib.Workspace(dev)
ids = torch.stack([torch.randint(0, nU, (N,), device=dev, generator=g), torch.randint(0, nI, (N,), device=dev, generator=g),
                   torch.randint(1, nI, (N,), device=dev, generator=g)])
steps = (N + B - 1) // B
losses = torch.zeros(steps, device=dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
lib = _lib.load()


def ev_time(fn, reps, flush_on=False):
    ms = []
    for _ in range(reps):
        if flush_on:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return float(np.median(ms)), float(np.mean(ms)), float(np.min(ms))


def phases(tr, rows):
    """Median ns of each phase over the given steps of a [steps, 8] stamp array."""
    t = tr[rows].astype(np.int64)
    d = {'bpr': t[:, 1] - t[:, 0], 'barrier1': t[:, 2] - t[:, 1], 'adam': t[:, 3] - t[:, 2], 'barrier2': t[:, 4] - t[:, 3]}
    d['step'] = np.diff(tr[:, 0].astype(np.int64))[rows[:-1]]
    return {k: float(np.median(v)) for k, v in d.items()}


# ---- 1. epoch launches ----
k = [0]
def epoch():
    _lib.bprmf_epoch(P, M, V, G, ids, B, nU, k[0], 1e-3, 1e-6, losses, ws)
    k[0] += steps
epoch(); torch.cuda.synchronize()
med, mean, mn = ev_time(epoch, 7)
out['epoch_launch'] = {'steps': steps, 'ms_median': med, 'ms_min': mn, 'us_per_step': med * 1e3 / steps}
trace = torch.zeros((steps, 8), dtype=torch.int64, device=dev)
ctr = torch.zeros((steps, 256, 4), dtype=torch.int64, device=dev)
lib.wr_debug_epoch_trace(trace.data_ptr(), ctr.data_ptr())
epoch(); torch.cuda.synchronize()
lib.wr_debug_epoch_trace(None, None)
tr = trace.cpu().numpy()
out['epoch_phases_ns'] = phases(tr, np.arange(10, steps - 2))
out['epoch_losses_finite'] = bool(torch.isfinite(losses).all())
sms = torch.cuda.get_device_properties(0).multi_processor_count
c = ctr.cpu().numpy().reshape(-1)[:steps * sms * 4].reshape(steps, sms, 4)[10:steps - 2].astype(np.int64)   # [steps, grid, 4]
out['epoch_barriers_ns'] = {
    'b1_arrival_skew': float(np.median(c[:, :, 0].max(1) - c[:, :, 0].min(1))),
    'b1_last_arrival_to_last_pass': float(np.median(c[:, :, 1].max(1) - c[:, :, 0].max(1))),
    'b1_last_arrival_to_first_pass': float(np.median(c[:, :, 1].min(1) - c[:, :, 0].max(1))),
    'b2_arrival_skew': float(np.median(c[:, :, 2].max(1) - c[:, :, 2].min(1))),
    'b2_last_arrival_to_last_pass': float(np.median(c[:, :, 3].max(1) - c[:, :, 2].max(1))),
    'bpr_phase_max_cta': float(np.median(c[1:, :, 0].max(1) - c[:-1, :, 3].min(1))),
    'adam_phase_max_cta': float(np.median(c[:, :, 2].max(1) - c[:, :, 1].min(1))),
    'adam_phase_median_cta': float(np.median(c[:, :, 2] - c[:, :, 1])),
    'bpr_phase_median_cta': float(np.median(c[1:, :, 0] - c[:-1, :, 3])),
    'b1_wait_median_cta': float(np.median(c[:, :, 1] - c[:, :, 0])), 'b2_wait_median_cta': float(np.median(c[:, :, 3] - c[:, :, 2]))}

if len(sys.argv) > 2 and sys.argv[2] == 'epoch-only':
    print(json.dumps(out, indent=1)); sys.exit(0)
# ---- 2. one launch per step ----
u, p_, n_ = ids[0, :B], ids[1, :B], ids[2, :B]
loss1 = torch.zeros(1, device=dev)
def step():
    k[0] += 1
    _lib.bprmf_step(P, M, V, G, u, p_, n_, nU, k[0], 1e-3, 1e-6, loss1, ws)
def old_step():
    k[0] += 1
    _lib.bprmf_step(P, M, V, G, u, p_.clone(), n_, nU, k[0], 1e-3, 1e-6, loss1, ws)     # unequal spacing: the L2-streamed form
tiny = [torch.zeros(4, device=dev) for _ in range(4)]
floor = lambda: _lib.adam_l2_sweep(*tiny, 1, 1e-3, 0.0)
for name, fn in (('floor_tiny_kernel', floor), ('resident_single_step', step)):
    fn(); torch.cuda.synchronize()
    out[name + '_us'] = {'warm': [x * 1e3 for x in ev_time(fn, 40)], 'flushed': [x * 1e3 for x in ev_time(fn, 40, True)]}
pc = p_.clone()
def old_step():  # noqa: F811
    k[0] += 1
    _lib.bprmf_step(P, M, V, G, u, pc, n_, nU, k[0], 1e-3, 1e-6, loss1, ws)
old_step(); torch.cuda.synchronize()
out['l2_streamed_single_step_us'] = {'warm': [x * 1e3 for x in ev_time(old_step, 40)],
                                     'flushed': [x * 1e3 for x in ev_time(old_step, 40, True)]}
# the phases of a flushed single-step launch
trace1 = torch.zeros((1, 8), dtype=torch.int64, device=dev)
lib.wr_debug_epoch_trace(trace1.data_ptr(), None)
ph = []
for _ in range(20):
    flush.zero_(); step(); torch.cuda.synchronize()
    t = trace1.cpu().numpy()[0]
    ph.append([t[1] - t[0], t[2] - t[1]])
lib.wr_debug_epoch_trace(None, None)
out['single_step_flushed_phases_ns'] = {'bpr': float(np.median([x[0] for x in ph])), 'barrier1': float(np.median([x[1] for x in ph]))}

# ---- 3. host-fed streaming ----
host_ids = ids[:, :64 * B].cpu()
pinned = [host_ids[:, i * B:(i + 1) * B].contiguous().pin_memory() for i in range(64)]
pageable = [host_ids[:, i * B:(i + 1) * B].contiguous() for i in range(64)]
ctx = _lib.BprmfContext(P, M, V, G, nU, 1e-3, 1e-6, ws)
cur = torch.cuda.current_stream()
t_adam = [k[0]]
def hstep(buf, wait=1):
    t_adam[0] += 1
    return ctx.step(buf.data_ptr(), B, t_adam[0], wait)
for i in range(20):
    hstep(pinned[i % 64])
for name, bufs, fl in (('pinned_wait1', pinned, False), ('pageable_wait1', pageable, False), ('pinned_wait1_flushed', pinned, True),
                       ('pageable_wait1_flushed', pageable, True)):
    lat = []
    for i in range(200 if not fl else 40):
        if fl:
            flush.zero_(); cur.synchronize()
        t0 = time.perf_counter()
        hstep(bufs[i % 64])
        lat.append(time.perf_counter() - t0)
    out['stream_' + name + '_us'] = {'median': float(np.median(lat)) * 1e6, 'mean': float(np.mean(lat)) * 1e6,
                                     'min': float(np.min(lat)) * 1e6}
for name, bufs in (('pinned', pinned), ('pageable', pageable)):
    n_push = 2000
    base = ctx.steps
    t0 = time.perf_counter()
    for i in range(n_push):
        hstep(bufs[i % 64], wait=0)
        if i >= 12:
            ctx.wait(base + i - 12, 1)
    for i in range(n_push - 12, n_push):
        ctx.wait(base + i, 1)
    dt = time.perf_counter() - t0
    out['stream_pipelined_' + name] = {'us_per_step': dt / n_push * 1e6, 'interactions_per_s': n_push * B / dt}
# GPU-side stamps of unpipelined steps (the kernel is relaunched so that first_step is known)
ctx.sync()
n_tr = 64
trace2 = torch.zeros((n_tr + 8, 8), dtype=torch.int64, device=dev)
lib.wr_debug_epoch_trace(trace2.data_ptr(), None)
host_t = []
for i in range(n_tr):
    t0 = time.perf_counter()
    hstep(pinned[i % 64])
    host_t.append(time.perf_counter() - t0)
ctx.sync()
lib.wr_debug_epoch_trace(None, None)
t2 = trace2.cpu().numpy()[:n_tr].astype(np.int64)
rows = np.arange(4, n_tr)
out['stream_unpipelined_gpu_side_ns'] = {
    'poller_seen_to_staged': float(np.median(t2[rows, 6] - t2[rows, 5])),
    'staged_to_step_start': float(np.median(t2[rows, 0] - t2[rows, 6])),
    'bpr': float(np.median(t2[rows, 1] - t2[rows, 0])), 'barrier1': float(np.median(t2[rows, 2] - t2[rows, 1])),
    'adam': float(np.median(t2[rows, 3] - t2[rows, 2])), 'barrier2': float(np.median(t2[rows, 4] - t2[rows, 3])),
    'barrier2_to_done_word': float(np.median(t2[rows, 7] - t2[rows, 4])),
    'poller_seen_to_done_word': float(np.median(t2[rows, 7] - t2[rows, 5])),
    'host_call_us': float(np.median(host_t[4:])) * 1e6}
ctx.close()
assert ws.status() == 0
print(json.dumps(out, indent=1))
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], 'w'), indent=1)
