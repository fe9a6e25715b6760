"""Per-kernel resource usage and SASS instruction mix of libcmarl_b200.so (runs without a GPU).

    python profiles/tools/resources.py > profiles/resources_r1.md

Sources: ``cuobjdump --dump-resource-usage`` (registers, stack = spills, static shared memory) and ``cuobjdump -sass``
(instruction count and the share of the classes that matter for this path: tcgen05 / UTC* tensor-core and TMEM traffic,
FFMA/FADD/FMUL, DFMA-class fp64, shared-memory and global-memory accesses, barriers).  Static counts: loops are
counted once."""
import re
import subprocess
import sys
from collections import Counter, OrderedDict
from pathlib import Path

LIB = Path(__file__).resolve().parents[2] / "cleanmarl_b200" / "libcmarl_b200.so"

CLASSES = OrderedDict([
    ("tensor (UTC*MMA)", r"^UTC\w*MMA"),
    ("TMEM ld/st (LDTM/STTM)", r"^(LDTM|STTM)"),
    ("tcgen05 other (UTC*)", r"^UTC"),
    ("fp32 FMA/ADD/MUL", r"^(FFMA|FADD|FMUL)"),
    ("fp64", r"^D(FMA|ADD|MUL|SETP)"),
    ("MUFU", r"^MUFU"),
    ("shared ld/st", r"^(LDS|STS|LDSM|ATOMS)"),
    ("global ld/st", r"^(LDG|STG|LD\.|ST\.|LD$|ST$|RED|ATOMG|ATOM\b)"),
    ("bulk copy (UBLKCP/TMA)", r"^(UBLKCP|UTMA)"),
    ("barriers (BAR/SYNCS)", r"^(BAR|SYNCS|WARPSYNC|ELECT)"),
    ("shuffles", r"^SHFL"),
])


def demangle(name):
    try:
        return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
    except OSError:
        return name


def short(name):
    d = demangle(name)
    d = d.replace("(anonymous namespace)::", "").replace("tcchain::", "").replace("chain::", "")
    d = re.sub(r"\(.*$", "", d)                         # drop the argument list
    d = re.sub(r"^void ", "", d)
    return d


def main():
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", str(LIB)], capture_output=True, text=True).stdout
    usage = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+)", res):
        usage[m.group(1)] = tuple(int(x) for x in m.groups()[1:])
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True).stdout
    mix, cur = {}, None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = mix.setdefault(m.group(1), Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["total"] += 1
            for cls, pat in CLASSES.items():
                if re.match(pat, op):
                    cur[cls] += 1
                    break
    print("# Static resources and instruction mix of `libcmarl_b200.so` (sm_100a, round 1)\n")
    print("`python profiles/tools/resources.py` — cuobjdump only, no GPU; loops are counted once. STACK > 0 would be spills or\n"
          "local arrays; dynamic shared memory (the chain / rollout / tbptt kernels) is set at launch and not listed here.\n")
    print("| kernel | regs | stack B | static smem B | SASS instr. | " + " | ".join(CLASSES) + " |")
    print("|---|---|---|---|---|" + "---|" * len(CLASSES))
    for name in sorted(usage, key=lambda n: -mix.get(n, Counter())["total"]):
        r, st, sh = usage[name]
        c = mix.get(name, Counter())
        print(f"| `{short(name)}` | {r} | {st} | {sh} | {c['total']} | " + " | ".join(str(c[k]) for k in CLASSES) + " |")


if __name__ == "__main__":
    sys.exit(main())
