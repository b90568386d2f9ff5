#!/usr/bin/env python
"""Print the headline numbers of bench.py JSON lines: python tools/benchline.py gpurun_out/bench_*.json"""
import json
import sys

for path in sys.argv[1:]:
    try:
        d = json.load(open(path))
    except Exception as e:      # noqa: BLE001
        print(path, "unreadable:", e)
        continue
    f = d.get("kernel_families_ms_per_step", {})
    print(f"{path}: {d['value']:.1f} steps/s, e2e {d['e2e']['value']:.1f}, roofline frac {d['roofline']['frac']:.3f}")
    print("   ", " ".join(f"{k}={v:.3f}" for k, v in f.items()))
