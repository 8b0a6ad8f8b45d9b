#!/usr/bin/env python
"""PCIe ceiling for the host-facing step: pinned D2H / H2D cudaMemcpy bandwidth (DMA engines) at the e2e step's sizes,
alone and both directions at once, to compare with what the zero-copy step kernel achieves writing 58 B/env itself."""
import json

import torch

n = 1 << 20
d2h_bytes, h2d_bytes = 58 * n, 12 * n
dev = torch.device("cuda:0")
src_d = torch.empty(d2h_bytes, dtype=torch.uint8, device=dev)
dst_h = torch.empty(d2h_bytes, dtype=torch.uint8).pin_memory()
src_h = torch.empty(h2d_bytes, dtype=torch.uint8).pin_memory()
dst_d = torch.empty(h2d_bytes, dtype=torch.uint8, device=dev)
s2 = torch.cuda.Stream()


def timed(fn, reps=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def both():
    s2.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s2):
        dst_d.copy_(src_h, non_blocking=True)
    dst_h.copy_(src_d, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s2)


ms_d2h = timed(lambda: dst_h.copy_(src_d, non_blocking=True))
ms_h2d = timed(lambda: dst_d.copy_(src_h, non_blocking=True))
ms_both = timed(both)
print(json.dumps({"d2h_bytes": d2h_bytes, "h2d_bytes": h2d_bytes, "d2h_ms": ms_d2h, "d2h_GBps": d2h_bytes / ms_d2h / 1e6,
                  "h2d_ms": ms_h2d, "h2d_GBps": h2d_bytes / ms_h2d / 1e6, "both_ms": ms_both,
                  "both_d2h_GBps": d2h_bytes / ms_both / 1e6}))
