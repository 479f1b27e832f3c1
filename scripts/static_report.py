"""Static facts of the ahead-of-time kernels (no GPU needed): registers, stack frame, spill bytes from `ptxas -v`; instruction mix from
`cuobjdump -sass` (FP64 arithmetic, MUFU, shared / local / global memory instructions, branches).

    python scripts/static_report.py > profiles/<round>_static_kernels.md

What it is for: the spill and instruction-mix statements of DESIGN.md section 4 can be checked without a device; ncu captures under
profiles/ give the dynamic side (executed counts, pipe utilisation)."""
import os, re, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "probnumdiffeq.jl_b200", "csrc")


def demangle(name):
    out = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    return re.sub(r"^void ", "", out)


def main():
    with tempfile.TemporaryDirectory() as tmp:
        obj = os.path.join(tmp, "pn_builtin.o")
        cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo", "-Xptxas=-v", "-I", CSRC, "-c",
               os.path.join(CSRC, "pn_builtin.cu"), "-o", obj]
        log = subprocess.run(cmd, capture_output=True, text=True)
        if log.returncode:
            sys.exit(log.stderr)
        res = {}
        for blk in log.stderr.split("Compiling entry function '")[1:]:
            name = blk.split("'")[0]
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", blk)
            r = re.search(r"Used (\d+) registers", blk)
            res[name] = dict(regs=int(r.group(1)), stack=int(m.group(1)), spill_st=int(m.group(2)), spill_ld=int(m.group(3)))
        sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
        cur = None
        for line in sass.splitlines():
            f = re.match(r"\s+Function : (\S+)", line)
            if f:
                cur = f.group(1)
                res.setdefault(cur, {}).update(n=0, fp64=0, mufu=0, shared=0, local=0, glob=0, branch=0)
                continue
            m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
            if not m or cur is None:
                continue
            op = m.group(1)
            c = res[cur]
            c["n"] += 1
            if op in ("DFMA", "DMUL", "DADD", "DSETP"):
                c["fp64"] += 1
            elif op == "MUFU":
                c["mufu"] += 1
            elif op in ("LDS", "STS", "LDSM"):
                c["shared"] += 1
            elif op in ("LDL", "STL"):
                c["local"] += 1
            elif op in ("LDG", "STG", "LD", "ST", "ATOMG", "RED"):
                c["glob"] += 1
            elif op in ("BRA", "BRX", "CALL", "RET", "EXIT"):
                c["branch"] += 1
    print("# Static facts of the ahead-of-time kernels (`scripts/static_report.py`: nvcc 12.9 `-Xptxas -v`, `cuobjdump -sass`, sm_100a)\n")
    print("Instruction counts are static (whole kernel, cold paths included), not executed counts.\n")
    print("| kernel | registers | stack B | spill st / ld B | instructions | FP64 | MUFU | LDS+STS | LDL+STL | global | branches |")
    print("|---|---|---|---|---|---|---|---|---|---|---|")
    for name in sorted(res, key=demangle):
        c = res[name]
        if "n" not in c or "regs" not in c:
            continue
        print(f"| `{demangle(name).split('(')[0]}` | {c['regs']} | {c['stack']} | {c['spill_st']} / {c['spill_ld']} | {c['n']} | {c['fp64']} "
              f"({100.0 * c['fp64'] / max(c['n'], 1):.0f} %) | {c['mufu']} | {c['shared']} | {c['local']} | {c['glob']} | {c['branch']} |")


if __name__ == "__main__":
    main()
