"""Opcode histogram of the shipped library (evidence for profiles/): python scripts/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections, re, subprocess, sys
LIB = "svn_icp_b200/lib/libsvnicp_b200.so"
elf = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True).stdout.strip()
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
ops, per_kernel, name = collections.Counter(), collections.defaultdict(collections.Counter), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and name:
        ops[m.group(1)] += 1
        per_kernel[name][m.group(1)] += 1
print(f"profiles/r02_sass_summary.txt -- cuobjdump -sass {LIB} (sm_100a only; round 2 final build, packed-fp32 k_gn)")
print(elf)
print("\nopcode histogram over all kernels (count opcode):")
print(";".join(f"{c} {o}" for o, c in ops.most_common(70)))
tc = sum(c for o, c in ops.items() if re.match(r"UTCMMA|UTMALDG|LDTM|HMMA|IMMA|DMMA|UTCHMMA|UTCIMMA", o))
print(f"\nmarkers: UBLKCP (1-D bulk TMA) {ops['UBLKCP']}, SYNCS (mbarrier) {ops['SYNCS']}, packed fp32 FFMA2 {ops['FFMA2']} / FADD2 {ops['FADD2']} / FMUL2 {ops['FMUL2']}, "
      f"scalar FFMA {ops['FFMA']}, DFMA {ops['DFMA']}, tensor-core / TMEM / tensor-map TMA (UTCMMA|UTMALDG|LDTM|HMMA|IMMA|DMMA): {tc} -- none, as north_star states "
      "(no dense contraction of useful size on this path)")
print(f"kernels: {len(per_kernel)}")
print("\nper kernel (instructions; FFMA2+FADD2+FMUL2; UBLKCP; LDS; DFMA):")
for k in sorted(per_kernel, key=lambda k: -sum(per_kernel[k].values())):
    c = per_kernel[k]
    d = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()
    d = re.sub(r"\((?:[^()]|\([^()]*\))*\)$", "", d).replace("(bool)", "")[:70]
    print(f"  {sum(c.values()):6d}  packed {c['FFMA2'] + c['FADD2'] + c['FMUL2']:5d}  UBLKCP {c['UBLKCP']:3d}  LDS {c['LDS']:4d}  DFMA {c['DFMA']:4d}  {d}")
