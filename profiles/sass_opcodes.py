"""profiles/sass_opcodes.txt: per-kernel counts of the Blackwell-specific SASS opcodes in the shipped library
(`cuobjdump -sass madtp_b200/libmadtp_b200.so`): UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor loads, LDTM / STTM =
tcgen05.ld / st (TMEM), UTCBAR = tcgen05.commit -> mbarrier, SYNCS = mbarrier ops, MUFU.EX2 / MUFU.TANH = SFU.
    python profiles/sass_opcodes.py"""
import collections
import hashlib
import re
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "madtp_b200" / "libmadtp_b200.so"
OPS = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "LDTM", "STTM", "UTCBAR", "SYNCS", "MUFU.EX2", "MUFU.TANH", "HMMA", "FFMA", "DFMA"]
out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True).stdout
counts, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", name).replace("madtp::", "")
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        for o in OPS:
            if op == o or op.startswith(o + "."):
                counts[cur][o] += 1
lines = [f"# SASS opcode counts per kernel, {LIB.name} sha256 {hashlib.sha256(LIB.read_bytes()).hexdigest()[:16]}",
         "kernel".ljust(64) + "".join(o.rjust(13) for o in OPS)]
tot = collections.Counter()
for k, c in counts.items():
    if not any(c[o] for o in OPS[:6]) and "kernel" not in k:
        continue
    lines.append(k[:63].ljust(64) + "".join(str(c[o]).rjust(13) for o in OPS))
    tot.update(c)
lines.append("TOTAL".ljust(64) + "".join(str(tot[o]).rjust(13) for o in OPS))
(ROOT / "profiles" / "sass_opcodes.txt").write_text("\n".join(lines) + "\n")
print("\n".join(lines[:12]), "\n...\n", lines[-1])
