"""Per-phase timing of the step kernels (debug build -DMD_PHASE_CLK, build/libmdstep_clk.so): where inside k_pre / k_dyn /
k_post a CTA spends its cycles, at the bench workload with L2 flushed before the step.  Development tool, not a bench.

    make -C metadrive_ped_b200/csrc OUT=../../build/libmdstep_clk.so EXTRA=-DMD_PHASE_CLK
    MD_LIB=build/libmdstep_clk.so python scripts/phase_clk.py [cfg2]
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from metadrive_ped_b200 import lib as mdlib  # noqa: E402
from metadrive_ped_b200.sim import BatchedSim  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
E = bench.WORKLOADS[wl]["envs"]
lib_, arrays, cfg = bench.build_world(E, 0, wl)
sim = BatchedSim(arrays, cfg, device=0)
if wl in ("cfg2", "cfg4"):
    b_arrays, b_cfg = lib_.build_world(list(range(len(lib_))), seed=0, **bench.bank_kw(lib_, wl))
    bank = BatchedSim(b_arrays, b_cfg, device=0)
    bank.reset()
    sim.reset()
    sim.attach_bank(bank, seed=1000)
dev = torch.device("cuda", 0)
A = sim.n_agents
act = torch.tensor([0.0, 1.0], device=dev).repeat(A, 1).contiguous()
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
flush2 = torch.ones(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
sim.reset()
for _ in range(150):
    sim.step(act, autoreset=True)
CTAS = 4096
tab = torch.zeros((4, CTAS, 16), dtype=torch.int64, device=dev)
h = mdlib.load()
h.md_debug_phase_clk.argtypes = [C.c_void_p]
assert h.md_debug_phase_clk(C.c_void_p(tab.data_ptr())) == 0
NAMES = {0: ("k_pre", ["phase1 sweep", "trigger", "compact", "idm: rows+map+routing lane", "idm: neighbour mask", "idm: front/back + lane change", "idm: steering+acc", "idm: actuate+store, other rounds"]), 1: ("k_dyn", ["row sweep + object staging", "pass list", "row loads + broad phase", "5 sub-steps", "store + work list"]),
         3: ("k_scan", ["prelude (first item)", "candidates", "items", "reduce+store", "later items"]),
         2: ("k_post", ["restore", "phase1", "2a r0: map_view+loc_ctx+cell", "2a r0: candidates", "2a r0: static", "2a r0: contacts", "2a later rounds+shfl", "phase2"])}
acc = {}
for rep in range(5):
    flush.zero_()
    if os.environ.get("FLUSH_READ"):
        flush2.sum()
    tab.zero_()
    torch.cuda.synchronize()
    sim.step(act, autoreset=True)
    torch.cuda.synchronize()
    t = tab.cpu().numpy()
    for k, (name, phases) in NAMES.items():
        rows = t[k]
        used = rows[:, 0] != 0
        r = rows[used]
        marks = [0, 1, 2, 3, 4, 13] if k in (1, 3) else ([0, 1, 2, 3, 4, 5, 6, 7, 13] if k == 0 else [0, 1, 2, 4, 5, 6, 7, 3, 13])
        if k == 2:
            if (r[:, 7] != 0).any():
                r = r[r[:, 7] != 0]
            else:
                marks = [0, 1, 2, 2, 2, 2, 2, 3, 13]
        d = np.stack([r[:, marks[i + 1]] - r[:, marks[i]] for i in range(len(marks) - 1)], 1).astype(np.float64)
        wall = (r[:, 12].max() - r[:, 15].min()) / 1e3      # globaltimer ns -> us, first CTA start to last CTA end
        start = (r[:, 15] - r[:, 15].min()) / 1e3
        acc.setdefault(k, []).append((d.mean(0), np.percentile(d.sum(1), [50, 90, 100]), wall, np.percentile(start, [50, 90, 100]), used.sum(),
                                      len(np.unique(r[:, 14]))))
r = t[0][t[0][:, 0] != 0]
dur = (r[:, 13] - r[:, 0]) / 1965.0
print("k_pre (last rep): n_act percentiles", np.percentile(r[:, 10], [10, 50, 90, 99, 100]), "T counts", {int(v): int((r[:, 11] == v).sum()) for v in np.unique(r[:, 11])})
for lo, hi in ((0, 32), (32, 48), (48, 64), (64, 80), (80, 200)):
    sel = (r[:, 10] >= lo) & (r[:, 10] < hi)
    if sel.any():
        print("   n_act in [%d, %d): %d CTAs, duration mean %.1f max %.1f us" % (lo, hi, sel.sum(), dur[sel].mean(), dur[sel].max()))
r = t[2][t[2][:, 8] != 0]
if len(r):
    for role, sel in (("agent", r[:, 14] == 1), ("traffic", r[:, 14] == 0)):
        if sel.any():
            q = r[sel]
            print("k_post phase 2, thread 0's vehicle = %s (%d CTAs): rows loaded %.1f us after the phase began, after_step %.1f, outputs %.1f, stores %.1f" % (
                role, sel.sum(), ((q[:, 8] - q[:, 3]) / 1965.0).mean(), ((q[:, 9] - q[:, 8]) / 1965.0).mean(), ((q[:, 10] - q[:, 9]) / 1965.0).mean(),
                ((q[:, 11] - q[:, 10]) / 1965.0).mean()))
r = t[1][t[1][:, 0] != 0]
dur = (r[:, 13] - r[:, 0]) / 1965.0
print("k_dyn (last rep): passes", {int(v): int((r[:, 8] == v).sum()) for v in np.unique(r[:, 8])}, "CTAs with contact candidates", int((r[:, 9] > 0).sum()),
      "alive per CTA percentiles", np.percentile(r[:, 10], [10, 50, 90, 100]))
for label, sel in (("1 pass, no candidates", (r[:, 8] == 1) & (r[:, 9] == 0)), ("1 pass, candidates", (r[:, 8] == 1) & (r[:, 9] > 0)), ("2+ passes", r[:, 8] > 1)):
    if sel.any():
        print("   %-24s %4d CTAs, duration mean %.1f p90 %.1f max %.1f us; active %.1f alive %.1f" % (label, sel.sum(), dur[sel].mean(), np.percentile(dur[sel], 90), dur[sel].max(), r[sel, 11].mean(), r[sel, 10].mean()))
for k, (name, phases) in NAMES.items():
    m = np.mean([a[0] for a in acc[k]], 0) / 1965.0   # cycles -> us at 1965 MHz
    tot = np.mean([a[1] for a in acc[k]], 0) / 1965.0
    print("%s: CTAs %d on %d SMs, kernel span %.1f us; CTA start p50/p90/max %s us; CTA duration p50/p90/max %s us" % (
        name, acc[k][0][4], acc[k][0][5], np.mean([a[2] for a in acc[k]]), np.round(np.mean([a[3] for a in acc[k]], 0), 1), np.round(tot, 1)))
    for p, v in zip(phases + ["tail"] * 8, m):
        print("    %-32s %7.1f us (mean over CTAs)" % (p, v))
