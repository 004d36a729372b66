"""CPU tool: SASS mnemonic counts per kernel of the built library (python tools/sass_mnemonics.py > profiles/rNN_sass_mnemonics.txt)."""
import collections
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else "auncel_b200/libauncel_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
want = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "UTCATOMSWS", "UTMALDG", "UTMASTG", "SYNCS", "ELECT", "FFMA", "FADD", "FMUL", "HMMA", "REDUX", "ATOMG", "RED"]
print("SASS evidence (cuobjdump -sass auncel_b200/libauncel_b200.so, sm_100a): instruction counts per kernel")
print("UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM = tcgen05.ld, UTCATOMSWS = tcgen05.alloc/dealloc,")
print("UTMALDG = cp.async.bulk.tensor (TMA load), SYNCS = mbarrier ops, ELECT = elect.sync; FFMA = 0 in the")
print("distance kernels is deliberate (-fmad=false: separately rounded mul / add, bit-exactness)\n")
cur, counts = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        if op in want:
            counts[cur][op] += 1
for k, c in counts.items():
    print(k)
    print("    " + "  ".join(f"{o}={c[o]}" for o in want if c[o]))
