"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum ... --csv`) per kernel family.
usage: python scripts/summarize_launches.py gpurun_out/TAG/launches.csv [launch_metrics.csv]"""
import csv, re, sys, collections

def rows(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    return list(csv.DictReader(lines))

def short(name):
    m = re.match(r"(?:void )?(?:otvm::)?(\w+)(<.*>)?\(", name)
    base = m.group(1) if m else name[:40]
    t = m.group(2) or "" if m else ""
    t = t.replace("__nv_bfloat16", "bf16").replace("(bool)", "").replace("(int)", "")
    return base + t

def main():
    path = sys.argv[1]
    per = collections.OrderedDict()
    launches = collections.OrderedDict()
    for r in rows(path):
        k = (r["ID"], short(r["Kernel Name"]), r["Grid Size"], r["Block Size"])
        launches.setdefault(k, {})[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    unit = {r["Metric Name"]: r["Metric Unit"] for r in rows(path)}
    tot = 0.0
    for (i, name, grid, block), m in launches.items():
        d = m.get("gpu__time_duration.sum", 0.0)
        if unit.get("gpu__time_duration.sum") in ("ns", "nsecond"): d /= 1e3
        elif unit.get("gpu__time_duration.sum") in ("ms", "msecond"): d *= 1e3
        e = per.setdefault(name, dict(n=0, us=0.0, dram=0.0, l2=0.0, tensor=0.0))
        e["n"] += 1; e["us"] += d; tot += d
        e["dram"] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
        e["l2"] += m.get("lts__t_bytes.sum", 0.0)
        e["tensor"] += m.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0) * d
    print(f"# {path}: {sum(e['n'] for e in per.values())} launches, {tot:.1f} us serialised device time; units: dram {unit.get('dram__bytes_read.sum')}, l2 {unit.get('lts__t_bytes.sum')}")
    print(f"{'kernel':70s} {'n':>4s} {'us':>9s} {'share':>6s} {'us/launch':>9s} {'dram/launch':>12s} {'tensor%':>8s}")
    for name, e in sorted(per.items(), key=lambda kv: -kv[1]["us"]):
        print(f"{name[:70]:70s} {e['n']:4d} {e['us']:9.1f} {e['us']/tot:6.1%} {e['us']/e['n']:9.2f} {e['dram']/e['n']:12.4g} {e['tensor']/max(e['us'],1e-9):8.1f}")

if __name__ == "__main__":
    main()
