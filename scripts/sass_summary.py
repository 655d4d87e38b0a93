"""Static evidence for the shipped library (no GPU needed): per kernel of sdflib_b200/libsdfb200.so the SASS instruction count,
the mnemonics that show what the kernel is built from (packed float32 FFMA2 / FMUL2 / FADD2, 1-D TMA bulk copies UBLKCP +
mbarrier SYNCS, warp reductions REDUX, 128-bit loads, float64 arithmetic, shared-memory accesses) and registers / spills from
the ptxas logs of the build (sdflib_b200/build/*.ptxas.log).

    python scripts/sass_summary.py > profiles/r2_sass_summary.md
"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "sdflib_b200", "libsdfb200.so")
KEYS = ["FFMA2", "FMUL2", "FADD2", "FFMA", "DFMA", "DADD", "DMUL", "MUFU", "LDG.E.128", "LDG", "STG", "LDS", "STS", "UBLKCP", "SYNCS", "REDUX",
        "SHFL", "VOTE", "ATOM", "RED", "BAR"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def short(name):
    name = re.sub(r"sdfb200::\(anonymous namespace\)::", "", name)
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"^void ", "", name)
    return name


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels.setdefault(cur, collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            c = kernels[cur]
            c["total"] += 1
            for k in KEYS:
                if op == k or op.startswith(k + ".") or (k == "LDG.E.128" and op.startswith("LDG.E.128")):
                    c[k] += 1
    regs = {}
    for log in glob.glob(os.path.join(ROOT, "sdflib_b200", "build", "*.ptxas.log")):
        name = None
        spill = ""
        for line in open(log):
            m = re.search(r"Compiling entry function '(\S+)'", line)
            if m:
                name = m.group(1)
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
            if m and name:
                spill = m.group(2)
            m = re.search(r"Used (\d+) registers", line)
            if m and name:
                regs[name] = (m.group(1), spill)
    dm = demangle(list(kernels))
    own = [(short(dm[k]), k) for k in kernels if "cub" not in dm[k]]
    print("# SASS summary of `sdflib_b200/libsdfb200.so` (sm_100a, `cuobjdump -sass`; `scripts/sass_summary.py`)\n")
    print("Instruction counts are static (per kernel image). `regs` / `spill B` from the ptxas logs of the same build. CUB's radix-sort kernels")
    print("(mesh ingestion) are left out. Columns: packed float32 (`FFMA2`+`FMUL2`+`FADD2`), scalar `FFMA`, float64 (`DFMA`+`DADD`+`DMUL`), `MUFU`,")
    print("128-bit global loads, all global loads / stores, shared loads / stores, 1-D TMA bulk copies (`UBLKCP`) and mbarrier ops (`SYNCS`), `REDUX`, `SHFL`, `VOTE`, atomics.\n")
    print("| kernel | instr | regs | spill B | f32x2 | FFMA | f64 | MUFU | LDG.128 | LDG | STG | LDS | STS | UBLKCP | SYNCS | REDUX | SHFL | VOTE | ATOM+RED |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    for name, k in sorted(own):
        c = kernels[k]
        r, sp = regs.get(k, ("", ""))
        row = [name, c["total"], r, sp, c["FFMA2"] + c["FMUL2"] + c["FADD2"], c["FFMA"], c["DFMA"] + c["DADD"] + c["DMUL"], c["MUFU"], c["LDG.E.128"], c["LDG"],
               c["STG"], c["LDS"], c["STS"], c["UBLKCP"], c["SYNCS"], c["REDUX"], c["SHFL"], c["VOTE"], c["ATOM"] + c["RED"]]
        print("| " + " | ".join(str(x) for x in row) + " |")
    tot = collections.Counter()
    for _, k in own:
        tot.update(kernels[k])
    print(f"\n{len(own)} kernels of the repo's own, {tot['total']} instructions; `UBLKCP` {tot['UBLKCP']}, `SYNCS` {tot['SYNCS']}, `REDUX` {tot['REDUX']}, "
          f"packed float32 {tot['FFMA2'] + tot['FMUL2'] + tot['FADD2']}; no `UTCMMA` / `UTMALDG` (no dense contraction, no 2-D tiles on this path).")


if __name__ == "__main__":
    main()
