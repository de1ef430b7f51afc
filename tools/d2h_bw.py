"""Device -> host bandwidth of all GPUs at once into one host buffer (what the getters of a team do)."""
import os, sys, time, threading
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastsk_b200.fastsk import pinned_empty
n = torch.cuda.device_count()
per = 2 << 30     # bytes per GPU
for kind in ("pinned_empty (fsk_host_alloc)", "torch pin_memory"):
    host = pinned_empty((n * per // 8,)) if kind.startswith("pinned_empty") else torch.empty(n * per // 8, dtype=torch.float64).pin_memory().numpy()
    ht = torch.from_numpy(host)
    devs = [torch.empty(per // 8, dtype=torch.float64, device=f"cuda:{i}") for i in range(n)]
    for which in ("all GPUs at once", "one GPU alone", "GPUs 0-3", "GPUs 4-7"):
        ids = list(range(n)) if which == "all GPUs at once" else [0] if which == "one GPU alone" else [i for i in range(n) if (i < 4) == (which == "GPUs 0-3")]
        if not ids:
            continue
        for rep in range(2):
            for i in ids: torch.cuda.synchronize(i)
            t0 = time.perf_counter()
            for i in ids:
                with torch.cuda.device(i):
                    ht[i * (per // 8):(i + 1) * (per // 8)].copy_(devs[i], non_blocking=True)
            for i in ids: torch.cuda.synchronize(i)
            dt = time.perf_counter() - t0
        print(f"{kind}: {which}: {len(ids) * per / dt / 1e9:.1f} GB/s", flush=True)
