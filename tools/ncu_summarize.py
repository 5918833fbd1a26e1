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
# launch order of one forward (backbone.forward, default precision): SA1..4: fps, [grid build], ball query(+group), chain;
# FP1..3: three_nn, fp_assemble, chain; FP4: three_nn, fp_assemble (rows -> image), chain (y2 = points2 @ W0[:c2]), chain (gather + layers 1..)
CHAIN_STAGE = {1: "layer1:mlp", 2: "layer2:mlp", 3: "layer3:mlp", 4: "layer4:mlp", 5: "fa_layer1:mlp", 6: "fa_layer2:mlp", 7: "fa_layer3:mlp",
               8: "fa_layer4:interpolate", 9: "fa_layer4:mlp"}
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
        stage = CHAIN_STAGE.get(count["chain"])
    elif "three_nn" in name:
        count["nn"] += 1; stage = "fa_layer%d:three_nn" % count["nn"]
    elif "fp_assemble" in name:
        count["asm"] += 1; stage = "fa_layer%d:interpolate" % count["asm"]
    if stage:
        rd = to_bytes(r[ci["dram__bytes_read.sum"]], units[ci["dram__bytes_read.sum"]])
        wr = to_bytes(r[ci["dram__bytes_write.sum"]], units[ci["dram__bytes_write.sum"]])
        dur = to_ms(r[ci["gpu__time_duration.sum"]], units[ci["gpu__time_duration.sum"]])
        t = traffic.setdefault(stage, {"dram_read_bytes": 0, "dram_write_bytes": 0, "traffic_bytes": 0, "ncu_duration_ms": 0.0, "kernel": ""})
        t["dram_read_bytes"] += int(rd); t["dram_write_bytes"] += int(wr); t["traffic_bytes"] += int(rd + wr)
        t["ncu_duration_ms"] = round(t["ncu_duration_ms"] + dur, 6)
        t["kernel"] = (t["kernel"] + " + " if t["kernel"] else "") + name[:60]
csv.writer(open(sys.argv[2], "w")).writerows(out)
json.dump(traffic, open(sys.argv[3], "w"), indent=1)
print("kernels", len(data), "stages", len(traffic))
