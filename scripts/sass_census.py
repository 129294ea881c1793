"""SASS instruction census of libmdstep.so per kernel (cuobjdump -sass): instruction count, the top mnemonics, and the ones
that show how memory is moved (LDG / STG / LDS / STS / ATOM / UBLKCP = cp.async.bulk / SYNCS = mbarrier / SHFL / BAR).
No tensor-core instruction is expected on this path (nothing is a dense contraction)."""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "metadrive_ped_b200/libmdstep.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
usage = {}
for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+)", res):
    usage[m.group(1)] = (int(m.group(2)), int(m.group(3)), int(m.group(4)))
kern, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); kern[cur] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        kern[cur][m.group(1).split(".")[0]] += 1
watch = ["LDG", "STG", "LDS", "STS", "LDL", "STL", "ATOM", "ATOMS", "ATOMG", "RED", "UBLKCP", "SYNCS", "SHFL", "BAR", "FFMA", "FMUL", "FADD", "MUFU",
         "HMMA", "UTCHMMA", "UTCQMMA"]
print("arch:", re.search(r"arch = (\S+)", out).group(1))
for name, c in kern.items():
    short = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.split("(")[0].strip()
    u = usage.get(name, (0, 0, 0))
    print("\n%s  (%d SASS instructions, %d registers, %d B stack, %d B static smem)" % (short, sum(c.values()), u[0], u[1], u[2]))
    print("   top: " + ", ".join("%s %d" % kv for kv in c.most_common(10)))
    print("   memory / sync / fp32: " + ", ".join("%s %d" % (k, c[k]) for k in watch if c[k]))
