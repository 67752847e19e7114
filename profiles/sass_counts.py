"""Opcode counts per kernel from `cuobjdump -sass` of the shipped library (the SASS evidence the
profiling recipe asks for: UTCHMMA / LDTM / STTM / UTCBAR prove tcgen05 + TMEM, RED the vector
reductions of the hash-grid scatter, MUFU the softplus epilogue, DFMA the fp64 filter).

    python profiles/sass_counts.py > profiles/r02_sass_counts.txt"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "deblur-e-nerf_b200", "lib", "libden_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
WATCH = ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTMALDG", "UTMASTG", "MUFU", "RED", "REDG", "ATOM", "ATOMG",
         "ATOMS", "DFMA", "DMUL", "DADD", "SYNCS", "NANOSLEEP", "FFMA", "FADD", "FMUL", "LDG", "STG", "LDS", "STS",
         "SHFL", "BAR", "F2FP", "HMMA", "LDL", "STL")
cur, counts, sizes = None, {}, {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur:
        counts[cur][m.group(1)] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)} (sm_100a): opcode counts per kernel\n")
for mangled, name in zip(counts, demangle):
    c = counts[mangled]
    total = sum(c.values())
    short = name.split("(")[0]
    print(f"## {short}   [{total} SASS instructions]")
    print("   " + "  ".join(f"{k} {c[k]}" for k in WATCH if c.get(k)))
