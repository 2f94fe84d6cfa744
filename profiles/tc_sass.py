"""SASS evidence that the tensor family uses tcgen05 / tensor memory: per kernel of fbp_tc.o, the counts of the
mnemonics B200_PROFILING.md lists (UTCHMMA = tcgen05.mma, STTM / LDTM = tcgen05.st / ld, UTCBAR = tcgen05.commit,
UTCATOMSWS = tcgen05.alloc / dealloc, SYNCS = mbarrier) plus FFMA2 (the weight-gradient warps) and SHFL.
Needs only cuobjdump (no GPU):   make -C fbpinns_b200/csrc && python profiles/tc_sass.py > profiles/r1f_tc_sass.md"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
obj = os.path.join(ROOT, "fbpinns_b200", "csrc", "build", "fbp_tc.o")
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
log = open(os.path.join(ROOT, "fbpinns_b200", "csrc", "build", "fbp_tc.ptxas.log")).read()
regs = dict(re.findall(r"Compiling entry function '(\S+)' for 'sm_100a'[\s\S]*?Used (\d+) registers", log))
print("# SASS mnemonic counts of the tensor-family kernels (`cuobjdump -sass fbpinns_b200/csrc/build/fbp_tc.o`, sm_100a)\n")
print("| kernel | registers | UTCHMMA | STTM | LDTM | UTCBAR | UTCATOMSWS | SYNCS | FFMA2 | SHFL | instructions |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n", 1)[0].strip()
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    ops = collections.Counter(re.findall(r"^\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P[0-9T]+\s+)?([A-Z0-9_]+)", f, flags=re.M))
    print(f"| `{dem[:90]}` | {regs.get(name, '?')} | {ops['UTCHMMA']} | {ops['STTM']} | {ops['LDTM']} | {ops['UTCBAR']} | "
          f"{ops['UTCATOMSWS']} | {ops['SYNCS']} | {ops['FFMA2']} | {ops['SHFL']} | {sum(ops.values())} |")
