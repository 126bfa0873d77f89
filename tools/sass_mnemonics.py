"""Per-kernel counts of the SASS mnemonics that prove a Blackwell-native kernel (B200_PROFILING.md: tcgen05.mma -> UTC*MMA,
tcgen05.ld/st -> LDTM/STTM, bulk async copies -> UBLKCP, mbarrier -> SYNCS, elect.sync -> ELECT; HMMA would be the legacy
mma.sync path). Runs anywhere (cuobjdump on the built library): `python tools/sass_mnemonics.py > profiles/sass_r1_mnemonics.txt`."""
import collections
import os
import re
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nsdp_b200 import _lib  # noqa: E402

KEYS = ("UTCHMMA", "UTCBAR", "UBLKCP", "LDTM", "STTM", "SYNCS", "ELECT", "HMMA", "FFMA", "UCGABAR_ARV", "REDUX")
txt = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
print(f"# {os.path.basename(_lib.LIB_PATH)}: SASS mnemonic counts per kernel (static instruction counts, sm_100a)")
print("# " + " ".join(KEYS))
rows = []
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n", 1)[0].strip()
    cnt = collections.Counter(m.group(1) for m in re.finditer(r"\b(" + "|".join(KEYS) + r")\b", f))
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    rows.append((re.sub(r"\(.*", "", dem).replace("void ", ""), cnt))
for dem, cnt in sorted(rows):
    print(dem[:100].ljust(102) + " ".join(f"{k}={cnt[k]}" for k in KEYS if cnt[k]))
