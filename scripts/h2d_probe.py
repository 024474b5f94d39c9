"""Host<->device copy ceilings of the box (one process, every visible GPU): what the end-to-end bfe path (pgc_eval_host: pinned host
buffer -> H2D -> kernels -> D2H) can reach at 1 / 2 / 4 / 8 GPUs.

For every GPU alone, for every pair (0, k), and for the sets {0,1}, {0..3}, {0..7}: aggregate GB/s of concurrent pinned-host ->
device copies (256 MiB each, 8 repetitions, one stream per device, enqueued from one thread), then the same for device -> host.
Two GPUs whose pair rate is about ONE GPU's rate share a PCIe uplink; a set whose rate stops growing has hit the host side
(root-complex or DRAM) ceiling.  Also: a STREAM-like host copy (numpy, 1 and N threads) for the DRAM figure the staging memcpy of
pageable input competes with.  Writes gpurun_out/h2d_probe.json."""
import json
import sys
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
n_dev = torch.cuda.device_count()
MB = 256
REPS = 8
host = [torch.empty(MB << 20, dtype=torch.uint8).pin_memory() for _ in range(n_dev)]
dev = [torch.empty(MB << 20, dtype=torch.uint8, device=f"cuda:{d}") for d in range(n_dev)]
streams = [torch.cuda.Stream(device=d) for d in range(n_dev)]
for h in host:
    h.fill_(1)


def rate(devs, h2d=True):
    for _ in range(2):  # warm-up + timed
        for d in devs:
            torch.cuda.synchronize(d)
        t0 = time.perf_counter()
        for _ in range(REPS):
            for d in devs:
                with torch.cuda.stream(streams[d]):
                    if h2d:
                        dev[d].copy_(host[d], non_blocking=True)
                    else:
                        host[d].copy_(dev[d], non_blocking=True)
        for d in devs:
            streams[d].synchronize()
        dt = time.perf_counter() - t0
    return len(devs) * REPS * (MB << 20) / dt / 1e9


out = {"devices": n_dev, "copy_mib": MB, "reps": REPS, "h2d_single_gbs": [rate([d]) for d in range(n_dev)],
       "d2h_single_gbs": [rate([d], False) for d in range(n_dev)]}
out["h2d_pair_with_0_gbs"] = {str(k): rate([0, k]) for k in range(1, n_dev)}
sets = [list(range(k)) for k in (1, 2, 4, 8) if k <= n_dev]
out["h2d_sets_gbs"] = {str(len(s)): rate(s) for s in sets}
out["d2h_sets_gbs"] = {str(len(s)): rate(s, False) for s in sets}
if n_dev >= 8:
    out["h2d_even_gpus_0_2_4_6_gbs"] = rate([0, 2, 4, 6])
    out["h2d_0_1_4_5_gbs"] = rate([0, 1, 4, 5])

# host DRAM: copy of a 1 GiB array, 1 thread and 8 threads (numpy releases the GIL in copyto)
a = np.ones(1 << 27)
b = np.empty_like(a)


def host_copy(threads):
    parts = np.array_split(np.arange(a.size), threads)
    with ThreadPoolExecutor(threads) as pool:
        for _ in range(2):
            t0 = time.perf_counter()
            list(pool.map(lambda p: np.copyto(b[p[0]:p[-1] + 1], a[p[0]:p[-1] + 1]), parts))
            dt = time.perf_counter() - t0
    return 2 * a.nbytes / dt / 1e9  # read + write


out["host_copy_read_plus_write_gbs"] = {str(t): host_copy(t) for t in (1, 8, 16)}
try:
    import subprocess
    out["topo"] = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=30).stdout[-3000:]
    out["lspci_tree"] = subprocess.run("lspci -tv 2>/dev/null | grep -i -B2 nvidia | head -80", shell=True, capture_output=True, text=True,
                                       timeout=30).stdout[-4000:]
except Exception as e:  # noqa: BLE001
    out["topo"] = str(e)
print(json.dumps({k: v for k, v in out.items() if k not in ("topo", "lspci_tree")}, indent=1))
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "h2d_probe.json").write_text(json.dumps(out, indent=1))
