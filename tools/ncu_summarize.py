"""Turn `ncu -i fwd.ncu-rep --page raw --csv` (a --set full capture of tools/one_forward.py) into the two committed
artefacts: a per-kernel summary csv and profiles/ncu_traffic.json (DRAM bytes per launch, keyed by bench.py stage name).
usage: python tools/ncu_summarize.py raw.csv out_summary.csv out_traffic.json"""
import csv, json, sys

COLS = ["launch__grid_size", "launch__block_size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem"]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def to_ms(v, unit):
    v = float(v.replace(",", ""))
    return v * {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}.get(unit, 1)


rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
ci = {h: i for i, h in enumerate(hdr)}
name_i = ci["Kernel Name"]
out = [["Kernel Name"] + COLS, [""] + [units[ci[c]] if c in ci else "" for c in COLS]]
traffic = {}
count = {"fps": 0, "bq": 0, "chain": 0, "nn": 0, "asm": 0}
for r in data:
    name = r[name_i]
    out.append([name[:60]] + [r[ci[c]] if c in ci else "" for c in COLS])
    stage = None
    if "fps_" in name:
        count["fps"] += 1; stage = "layer%d:fps" % count["fps"]
    elif "ballquery" in name:
        count["bq"] += 1; stage = "layer%d:ballquery_group" % count["bq"]
    elif "mlp_chain" in name:
        count["chain"] += 1
        stage = ("layer%d:mlp" % count["chain"]) if count["chain"] <= 4 else ("fa_layer%d:mlp" % (count["chain"] - 4))
    elif "three_nn" in name:
        count["nn"] += 1; stage = "fa_layer%d:three_nn" % count["nn"]
    elif "fp_assemble" in name:
        count["asm"] += 1; stage = "fa_layer%d:interpolate" % count["asm"]
    if stage:
        rd = to_bytes(r[ci["dram__bytes_read.sum"]], units[ci["dram__bytes_read.sum"]])
        wr = to_bytes(r[ci["dram__bytes_write.sum"]], units[ci["dram__bytes_write.sum"]])
        traffic[stage] = {"dram_read_bytes": int(rd), "dram_write_bytes": int(wr), "traffic_bytes": int(rd + wr),
                          "ncu_duration_ms": round(to_ms(r[ci["gpu__time_duration.sum"]], units[ci["gpu__time_duration.sum"]]), 6),
                          "kernel": name[:60]}
csv.writer(open(sys.argv[2], "w")).writerows(out)
json.dump(traffic, open(sys.argv[3], "w"), indent=1)
print("kernels", len(data), "stages", len(traffic))
