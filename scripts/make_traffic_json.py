"""profiles/ncu_traffic.json from an ncu per-launch metrics pass of ONE eager frame (scripts/gpu_round.sh step `ncu`):
DRAM bytes (read + write) per kernel family per frame, read by bench.py for `roofline.traffic`.
usage: python scripts/make_traffic_json.py gpurun_out/TAG/launch_metrics.csv TAG"""
import csv, json, sys, collections
path, tag = sys.argv[1], sys.argv[2]
lines = [l for l in open(path) if not l.startswith("==")]
fam = collections.defaultdict(lambda: dict(launches=0, dram_bytes=0.0, us=0.0))
seen = set()
for r in csv.DictReader(lines):
    n = r["Kernel Name"]
    f = ("conv_tcgen05" if "conv_tc_" in n else "memory_read" if "memory_read_tc" in n else "memory_read_combine" if "combine" in n
         else "gn_apply" if "gn_apply" in n else "conv_ffma" if ("conv_simt" in n or "smallm" in n) else "splitk_finish" if "splitk" in n
         else "upsample" if "upsample" in n else "other")
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    if r["Metric Name"] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        fam[f]["dram_bytes"] += v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    elif r["Metric Name"] == "gpu__time_duration.sum":
        fam[f]["us"] += v * {"ns": 1e-3, "us": 1, "ms": 1e3, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3}.get(u, 1)
        if r["ID"] not in seen:
            seen.add(r["ID"]); fam[f]["launches"] += 1
prec = sys.argv[3] if len(sys.argv) > 3 else "bf16x2"
out = {"precision": prec, "source": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum of one eager steady-state 512x512 T=8 frame ({tag}); "
                 "cold-cache, serialised launches", "per_frame": {k: {kk: round(vv, 1) for kk, vv in v.items()} for k, v in fam.items()}}
json.dump(out, open("profiles/ncu_traffic.json", "w"), indent=1)
print(json.dumps(out, indent=1))
