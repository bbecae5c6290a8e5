"""Instruction histogram of the specialised step kernel from the built library (cuobjdump -sass): opcode counts, share of
fp64 / conversions / shared-memory traffic, code bytes.  No GPU needed.
    python scripts/sass_hist.py [kernel-substring] > profiles/r02_sass_histogram.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "tiny-differentiable-simulator_b200", "libtds_b200.so")
want = sys.argv[1] if len(sys.argv) > 1 else "tds_step_spec_kernelI11SpecLaikagofdfLi1ELi1E"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
name, hist, n = None, collections.Counter(), 0
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        continue
    if name is None or want not in name:
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m:
        hist[m.group(1)] += 1
        n += 1
print(f"kernel {want}: {n} instructions = {n * 16 / 1024:.1f} KB of code (instruction caches: 32 KB L1.5, see r02_icache_probe.txt)")
groups = {"fp32 arithmetic": ("FFMA", "FMUL", "FADD", "FMNMX", "FSEL", "FSETP", "MUFU", "FFMA2"), "fp64 arithmetic": ("DFMA", "DMUL", "DADD", "DSETP"),
          "conversions": ("F2F", "F2I", "I2F", "I2FP", "FRND"), "shared memory": ("LDS", "STS", "LDSM"), "global memory": ("LDG", "STG"),
          "local memory (spills)": ("LDL", "STL"), "control": ("BRA", "BSSY", "BSYNC", "BAR", "CALL", "RET", "EXIT", "WARPSYNC", "BREAK")}
for g, ops in groups.items():
    c = sum(hist[o] for o in ops)
    print(f"  {g:24s} {c:6d}  {100.0 * c / max(n, 1):5.1f} %")
print("  opcodes:")
for op, c in hist.most_common(40):
    print(f"    {op:10s} {c:6d}")
