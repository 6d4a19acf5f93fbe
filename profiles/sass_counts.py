"""Per-kernel SASS instruction counts of libast_b200.so (proof of tcgen05 / TMA / TMEM use).
usage: python profiles/sass_counts.py > profiles/r02_sass_counts.txt"""
import re
import subprocess
import sys

LIB = sys.argv[1] if len(sys.argv) > 1 else "artist_style_transfer_b200/libast_b200.so"
OPS = ["UTCHMMA", "UTCQMMA", "LDTM", "UTMALDG", "UBLKCP", "UTCBAR", "SYNCS", "RED.E", "ATOMG", "DADD", "HMMA", "FFMA"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
blocks = re.split(r"\n\s*Function : \S+\n", sass)[1:]
print("# cuobjdump -sass artist_style_transfer_b200/libast_b200.so | per-kernel instruction counts (sm_100a)")
print("# UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor (TMA), UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit,")
print("# SYNCS = mbarrier ops, RED.E/ATOMG = global reductions / atomics, DADD = fp64 adds (InstanceNorm running sums), FFMA = fp32 FMA")
print(f"{'kernel':90s} " + " ".join(f"{o:>8s}" for o in OPS))
for name, blk in zip(names, blocks):
    short = re.sub(r"\(.*", "", name).replace("void ", "")
    counts = [len(re.findall(r"\b" + re.escape(o) + r"\b", blk)) for o in OPS]
    print(f"{short[:90]:90s} " + " ".join(f"{c:8d}" for c in counts))
