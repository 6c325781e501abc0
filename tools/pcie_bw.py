#!/usr/bin/env python3
"""Host<->device copy bandwidth of this box for the bench's transfer sizes (pinned memory, one stream each way, and both at once)."""
import json, time, torch
dev = torch.device("cuda:0")
n_in, n_out = 1024 * 1241 * 376, 1024 * 2048 * 60
h_in = torch.empty(n_in, dtype=torch.uint8).pin_memory(); d_in = torch.empty(n_in, dtype=torch.uint8, device=dev)
h_out = torch.empty(n_out, dtype=torch.uint8).pin_memory(); d_out = torch.empty(n_out, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=10):
    torch.cuda.synchronize(); t0 = time.time()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize(); return (time.time() - t0) / reps
run(True, True, 2)
a, b, c = run(True, False), run(False, True), run(True, True)
print(json.dumps({"h2d_GBps": n_in / a / 1e9, "d2h_GBps": n_out / b / 1e9, "both_ms": c * 1e3, "h2d_ms": a * 1e3, "d2h_ms": b * 1e3,
                  "both_h2d_GBps": n_in / c / 1e9}))
