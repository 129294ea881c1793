"""Top source lines of one kernel by warp-stall samples (from `ncu --page source --print-source cuda,sass --csv`)."""
import csv, subprocess, sys
rep, kernel = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      "regex:" + kernel], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = ""
agg = {}
seen_kernel = 0
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) == 2 and r[0] == "Function Name":
        seen_kernel += 1
        continue
    if len(r) > 8 and r[0].isdigit() and r[2] == "-":  # a source line row (aggregated over its SASS)
        if seen_kernel > 2:  # only the first launch of this kernel in the report (file sections repeat per launch)
            pass
        key = (cur_file, int(r[0]))
        samples, inst, tinst = int(r[4] or 0), int(r[7] or 0), int(r[8] or 0)
        a = agg.setdefault(key, [0, 0, 0, r[1]])
        a[0] += samples; a[1] += inst; a[2] += tinst
tot = sum(a[0] for a in agg.values()) or 1
print("total samples", tot)
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    print("%5.1f%%  inst %9d  thr/inst %5.1f  %s:%d  %s" % (100.0 * a[0] / tot, a[1], a[2] / max(a[1], 1), f, ln, a[3].strip()[:110]))
