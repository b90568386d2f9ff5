#!/usr/bin/env python
"""L2 -> shared-memory streaming ceilings on this GPU (csrc/streambench.cu): bulk async copies, 8 x 16 KB in flight per CTA,
one CTA per SM.  The deep levels of the network are bound by this stream (weights and replicated activation tiles), so
this is the roofline they are held against."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointsecguard_b200 import _lib as L

dev = torch.device("cuda", 0)
sms = torch.cuda.get_device_properties(dev).multi_processor_count
REGION = 512 << 10                      # bytes per CTA / cluster (L2-resident: 148 x 512 KB = 74 MB)
PASSES = 8
buf = torch.randint(0, 255, (sms * REGION,), dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream


def run(mode, cluster, ctas):
    L.psg_debug_l2_stream(buf.data_ptr(), REGION, mode, cluster, 2, ctas, st)      # warm the L2
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); L.psg_debug_l2_stream(buf.data_ptr(), REGION, mode, cluster, PASSES, ctas, st); e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    per_cta = REGION * PASSES / (best / 1e3) / 1e9
    return {"ms": round(best, 4), "GBps_delivered_per_SM": round(per_cta, 1), "TBps_delivered_total": round(per_cta * ctas / 1e3, 2),
            "TBps_from_L2": round(per_cta * ctas / cluster / 1e3, 2)}


out = {"sms": sms, "region_bytes": REGION, "passes": PASSES}
# in-flight depth and copy size: latency-bound or throughput-bound?
sweep = {}
for stages, sb in ((2, 16384), (4, 16384), (8, 16384), (12, 16384), (4, 4096), (16, 4096), (4, 32768), (6, 32768), (16, 2048), (16, 8192)):
    L.psg_set_option(b"stream_stage_bytes", 1024); L.psg_set_option(b"stream_stages", stages); L.psg_set_option(b"stream_stage_bytes", sb)
    sweep[f"{stages}x{sb // 1024}KB"] = {"one": run(0, 1, 1)["GBps_delivered_per_SM"], "all": run(0, 1, sms)["GBps_delivered_per_SM"],
                                          "all_same": run(1, 1, sms)["GBps_delivered_per_SM"]}
out["in_flight_sweep_GBps_per_SM"] = sweep
# several independent rings per CTA (one issuing thread each): is the per-copy cost in the issuing thread or in the copy engine?
rings = {}
for nr, stages, sb in ((1, 4, 16384), (2, 4, 16384), (4, 2, 16384), (8, 1, 16384), (4, 4, 8192), (8, 4, 4096), (8, 2, 8192), (4, 1, 32768), (2, 2, 32768)):
    L.psg_set_option(b"stream_stage_bytes", 1024); L.psg_set_option(b"stream_stages", stages); L.psg_set_option(b"stream_stage_bytes", sb)
    L.psg_set_option(b"stream_rings", nr)
    rings[f"{nr}rings_{stages}x{sb // 1024}KB"] = {"one": run(0, 1, 1)["GBps_delivered_per_SM"], "all": run(0, 1, sms)["GBps_delivered_per_SM"],
                                                  "all_same": run(1, 1, sms)["GBps_delivered_per_SM"]}
L.psg_set_option(b"stream_rings", 1)
out["rings_GBps_per_SM"] = rings
L.psg_set_option(b"stream_stage_bytes", 1024); L.psg_set_option(b"stream_stages", 8); L.psg_set_option(b"stream_stage_bytes", 16384)
full = sms // 8 * 8
out["one_cta_alone"] = run(0, 1, 1)
out["8_ctas_distinct"] = run(0, 1, 8)
out["32_ctas_distinct"] = run(0, 1, 32)
out["all_sms_distinct_regions"] = run(0, 1, sms)
out["all_sms_same_region"] = run(1, 1, sms)
for cs in (2, 4, 8):
    out[f"all_sms_multicast_cluster{cs}"] = run(0, cs, full)
    out[f"all_sms_multicast_cluster{cs}_same_region"] = run(1, cs, full)
print(json.dumps(out, indent=1))
