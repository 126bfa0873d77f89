"""Per-kernel registers / static shared memory / local-memory (spill) bytes of the built library, from
`cuobjdump --dump-resource-usage` (the numbers `-Xptxas -v` prints at build time). Runs anywhere:
`python tools/resource_usage.py > profiles/resources_r2.txt`. A non-zero LOCAL or STACK on a hot kernel means spills."""
import os
import re
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nsdp_b200 import _lib  # noqa: E402

txt = subprocess.run(["cuobjdump", "--dump-resource-usage", _lib.LIB_PATH], capture_output=True, text=True).stdout
rows = []
for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", txt):
    dem = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
    rows.append((re.sub(r"\(.*", "", dem).replace("void ", ""), *map(int, m.groups()[1:])))
print(f"# {os.path.basename(_lib.LIB_PATH)}: resource usage per kernel (sm_100a); dynamic shared memory is set by the launchers")
print(f"# {'kernel':<98} {'REG':>4} {'STACK':>6} {'SHARED':>7} {'LOCAL':>6}")
for name, reg, stack, shared, local in sorted(rows):
    print(f"{name[:100]:<100} {reg:>4} {stack:>6} {shared:>7} {local:>6}")
print(f"# {len(rows)} kernels; {sum(1 for r in rows if r[4])} with LOCAL (spill) bytes; "
      f"{sum(1 for r in rows if r[2])} with a stack frame (largest {max(r[2] for r in rows)} B)")
