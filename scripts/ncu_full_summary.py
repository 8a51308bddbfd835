"""One line per captured kernel from `ncu -i X.ncu-rep --page raw --csv` exports (scripts/gpu_round.sh step `ncufull`).
usage: python scripts/ncu_full_summary.py TAG gpurun_out/TAG/read_full.raw.csv gpurun_out/TAG/conv_full.raw.csv ..."""
import csv, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor"]
tag = sys.argv[1]
for path in sys.argv[2:]:
    rows = list(csv.reader(open(path)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"== {path.split('/')[-1]}  [ncu --set full --clock-control none, {tag}]")
    for r in body:
        parts = [f"{r[col['Kernel Name']][:64]}", f"grid={r[col['Grid Size']]}", f"block={r[col['Block Size']]}"]
        for k in KEYS:
            if k in col and r[col[k]] != "":
                parts.append(f"{k.split('.')[0].replace('sm__', '').replace('launch__', '')}={r[col[k]]}{units[col[k]]}")
        print("  -- " + "; ".join(parts))
