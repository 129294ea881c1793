"""Per-kernel DRAM bytes and FP32 flops per launch from an `ncu --set full` capture of `bench.py` (scripts/gpu_profile.sh)
-> profiles/r02_ncu_counters.json, which bench.py reads for `roofline.traffic` and `roofline_fp32`.

flops = FADD + FMUL + 2 x FFMA thread instructions (smsp__sass_thread_inst_executed_op_*_pred_on, reported per elapsed
cycle, times the elapsed SMSP cycles).  bench.py's "k_post" slot times k_scan + k_post, "k_reset" = k_restore_bank."""
import csv, json, subprocess, sys
rep = sys.argv[1]
out_path = sys.argv[2] if len(sys.argv) > 2 else "profiles/r02_ncu_counters.json"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[0]
col = {n: i for i, n in enumerate(h)}
def f(r, name):
    v = r[col[name]].replace(",", "")
    return float(v) if v not in ("", "n/a") else 0.0
units = rows[1]
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
SLOT = {"k_pre": "k_pre", "k_dyn": "k_dyn", "k_scan": "k_post", "k_post": "k_post", "k_restore_bank": "k_reset", "k_restore_post": "k_reset",
        "k_lidar": "k_lidar"}
acc, seen = {}, {}
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    short = name.replace("void ", "").split("(")[0].split("<")[0]
    if short not in SLOT or seen.get(short):
        continue            # the first captured launch of each kernel = the full-batch step
    seen[short] = True
    dram = sum(f(r, "dram__bytes_%s.sum" % k) * scale[units[col["dram__bytes_%s.sum" % k]]] for k in ("read", "write"))
    cyc = f(r, "smsp__cycles_elapsed.avg")
    flops = cyc * (f(r, "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum.per_cycle_elapsed")
                   + f(r, "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum.per_cycle_elapsed")
                   + 2.0 * f(r, "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed"))
    a = acc.setdefault(SLOT[short], {"dram_bytes": 0.0, "fp32_flops": 0.0, "kernels": []})
    a["dram_bytes"] += dram; a["fp32_flops"] += flops
    a["kernels"].append({"name": short, "duration_us": f(r, "gpu__time_duration.sum") * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3}.get(units[col["gpu__time_duration.sum"]], 1.0),
                         "dram_bytes": dram, "fp32_flops": flops, "grid": r[col["launch__grid_size"]], "block": r[col["launch__block_size"]],
                         "registers": r[col["launch__registers_per_thread"]]})
acc["_source"] = "ncu --set full --clock-control none of `python bench.py --steps 3 --warmup 3 --burnin 60` (scripts/gpu_profile.sh, %s)" % rep
json.dump(acc, open(out_path, "w"), indent=1)
for k, v in acc.items():
    if k != "_source":
        print(k, "dram %.2f MB" % (v["dram_bytes"] / 1e6), "fp32 %.1f MFLOP" % (v["fp32_flops"] / 1e6), [(x["name"], round(x["duration_us"], 1)) for x in v["kernels"]])
