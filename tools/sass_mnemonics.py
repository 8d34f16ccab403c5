#!/usr/bin/env python
"""SASS mnemonic counts per kernel of libznippy_cuda.so (what profiles/*_sass_mnemonics.txt hold):
`python tools/sass_mnemonics.py > profiles/r2b_sass_mnemonics.txt`.  Runs on the build machine (cuobjdump, no GPU)."""
import collections, os, re, subprocess, sys
SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "znippy_b200", "libznippy_cuda.so")
WATCH = ["UBLKCP", "UTMACMDFLUSH", "LDGSTS", "LDGDEPBAR", "DEPBAR", "CREDUX", "REDUX", "SHF", "PRMT", "LOP3", "IMAD", "BAR", "ATOMS", "LDS", "STS", "LDG",
         "STG", "MEMBAR", "SHFL", "VOTE", "FLO", "POPC", "BREV", "MATCH", "SYNCS", "HMMA", "UTCMMA", "UTMALDG"]
out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True, text=True).stdout.splitlines()
print("# SASS mnemonic counts per kernel of libznippy_cuda.so (cuobjdump -sass, sm_100a).")
print("# What to look for: UBLKCP/UTMACMDFLUSH = TMA bulk stores (cp.async.bulk) in the pattern decoders; LDGSTS (+ LDGDEPBAR/DEPBAR) = cp.async: staging in the")
print("# BLAKE3 kernels and the bit-stream ring of k_zseq1; CREDUX/REDUX = redux.sync in the compressor's lazy match pick and the exec kernels;")
print("# no HMMA/UTCMMA anywhere: this path has no tensor-core work.")
print("kernel | total | " + " ".join(WATCH))
cur, cnt, tot, k = None, None, 0, 0
def flush():
    if cur is not None:
        print(f"{cur} | {tot} | " + " ".join(f"{m}={cnt[m]}" for m in WATCH if cnt[m]))
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        flush()
        cur = re.sub(r"\(.*", "", names[k]); k += 1
        cnt, tot = collections.Counter(), 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
    if m and cur is not None:
        tot += 1
        cnt[m.group(1)] += 1
flush()
