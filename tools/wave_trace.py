#!/usr/bin/env python
"""Summarises the per-CTA, per-pass timestamps the persistent wavefront kernel writes when PCD_WAVE_TRACE=<prefix> is
set (csrc/sor_tiled.cu: [wait begin, pass begin, pass end, published] in globaltimer ns for every CTA and pass of the
LAST launch of a slab; dumped by pcd_slab_destroy as <prefix>_row<row0>.bin).

    python tools/wave_trace.py <prefix>_row*.bin [--json out.json] [--strips N]   (N = ceil(W / 504): adds per-chunk-row means)

Per slab: pass period (start-to-start of consecutive passes, median over CTAs), the share of a pass a CTA spends waiting
for its dependencies (neighbouring CTAs / the neighbouring GPU's flags), computing, and publishing, and the same split
for the CTAs next to a slab edge."""
import json
import sys

import numpy as np


def load(path):
    raw = np.fromfile(path, dtype=np.uint8)
    npass, maxc = np.frombuffer(raw[:8].tobytes(), dtype=np.int32)
    t = np.frombuffer(raw[8:].tobytes(), dtype=np.uint64).reshape(npass, maxc, 4).astype(np.int64)
    used = t[0, :, 1] > 0
    return t[:, used, :]


def summarise(t, strips=0):
    npass, ncta, _ = t.shape
    t0 = t[:, :, 0].min()
    wait = (t[:, :, 1] - t[:, :, 0]) / 1e3           # us
    comp = (t[:, :, 2] - t[:, :, 1]) / 1e3
    pub = (t[:, :, 3] - t[:, :, 2]) / 1e3
    period = np.diff(t[:, :, 1], axis=0) / 1e3 if npass > 1 else np.zeros((1, ncta))
    body = slice(2, None) if npass > 4 else slice(0, None)     # skip the ramp-up passes
    span = (t[-1, :, 3].max() - t[0, :, 0].min()) / 1e3
    extra = {}
    if strips and ncta % strips == 0:
        nby = ncta // strips
        extra = {"compute_us_by_chunk_row": [round(float(x), 1) for x in comp[body].mean(axis=0).reshape(nby, strips).mean(axis=1)],
                 "wait_us_by_chunk_row": [round(float(x), 1) for x in wait[body].mean(axis=0).reshape(nby, strips).mean(axis=1)]}
    return {**extra, "passes": int(npass), "ctas": int(ncta), "launch_us": float(span), "us_per_pass": float(span / npass),
            "pass_period_us_median": float(np.median(period[body])) if npass > 1 else None,
            "wait_us_mean": float(wait[body].mean()), "wait_us_p95": float(np.percentile(wait[body], 95)),
            "compute_us_mean": float(comp[body].mean()), "compute_us_max": float(comp[body].max()), "compute_us_min": float(comp[body].min()),
            "publish_us_mean": float(pub[body].mean()),
            "wait_share": float(wait[body].sum() / (wait[body].sum() + comp[body].sum() + pub[body].sum())),
            "first_cta_start_spread_us": float((t[0, :, 0].max() - t0) / 1e3),
            "slowest_cta_compute_us": float(comp[body].mean(axis=0).max()), "fastest_cta_compute_us": float(comp[body].mean(axis=0).min())}


if __name__ == "__main__":
    paths = [a for a in sys.argv[1:] if a.endswith(".bin")]
    out = {}
    for p in sorted(paths):
        out[p.split("/")[-1]] = summarise(load(p), int(sys.argv[sys.argv.index("--strips") + 1]) if "--strips" in sys.argv else 0)
    if "--json" in sys.argv:
        with open(sys.argv[sys.argv.index("--json") + 1], "w") as f:
            json.dump(out, f, indent=1)
    for k, v in out.items():
        print(k, json.dumps(v))
