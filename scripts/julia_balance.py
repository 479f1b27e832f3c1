"""Structural sanity check for the Julia host sources on a box without Julia: every block opener (module, struct, function, if,
for, while, try, begin, let, do, quote, macro) has its `end`, and parentheses / brackets / braces balance -- per file, ignoring
strings and comments.  Not a parser: it catches the class of error VERDICT r1 found (a `module` closed before its helpers), not
type errors.   python scripts/julia_balance.py julia/ProbNumDiffEqB200/src/*.jl"""
import re
import sys

OPEN = {"module", "baremodule", "struct", "function", "if", "for", "while", "try", "begin", "let", "do", "quote", "macro"}


def strip(src: str) -> str:
    out, i, n = [], 0, len(src)
    while i < n:
        c = src[i]
        if src.startswith('"""', i):
            j = src.find('"""', i + 3)
            j = n if j < 0 else j + 3
            out.append(" " * 1 + "\n" * src[i:j].count("\n"))
            i = j
        elif c == '"':
            j = i + 1
            while j < n and src[j] != '"':
                j += 2 if src[j] == "\\" else 1
            out.append('""')
            i = j + 1
        elif src.startswith("#=", i):
            j = src.find("=#", i + 2)
            i = n if j < 0 else j + 2
        elif c == "#":
            j = src.find("\n", i)
            i = n if j < 0 else j
        elif c == "'" and i + 2 < n and (src[i + 2] == "'" or (src[i + 1] == "\\" and i + 3 < n and src[i + 3] == "'")) and (i == 0 or not (src[i - 1].isalnum() or src[i - 1] in ")]_")):
            i += 3 if src[i + 2] == "'" else 4
            out.append("' '")
        else:
            out.append(c)
            i += 1
    return "".join(out)


def check(path: str) -> int:
    s = strip(open(path).read())
    depth, pairs, errs = 0, {"(": 0, "[": 0, "{": 0}, 0
    close = {")": "(", "]": "[", "}": "{"}
    for ln, line in enumerate(s.split("\n"), 1):
        nest = 0   # inside [] an `end` is an index, inside () `if`/`for` are generators / ternaries
        for tok in re.finditer(r"[A-Za-z_@][A-Za-z_0-9!]*|[()\[\]{}]", line):
            t = tok.group(0)
            if t in pairs:
                pairs[t] += 1
                nest += 1
            elif t in close:
                pairs[close[t]] -= 1
                nest -= 1
            elif nest > 0 or sum(pairs.values()) > 0:
                continue
            elif t in OPEN:
                if t == "struct" and line[:tok.start()].rstrip().endswith("mutable"):
                    pass
                before = line[:tok.start()].rstrip()
                if t in ("if", "for", "while") and before and not before.endswith(("begin", "else", "=", "&&", "||", ";", "(", "do", "end")) and not re.match(r"^\s*(@static)?\s*$", before):
                    continue   # trailing `x for x in ...` generators are handled by the paren rule; `a && if` does not occur here
                depth += 1
            elif t == "end":
                depth -= 1
                if depth < 0:
                    print(f"{path}:{ln}: unmatched end")
                    errs += 1
                    depth = 0
    if depth != 0:
        print(f"{path}: {depth} block(s) not closed")
        errs += 1
    for k, v in pairs.items():
        if v != 0:
            print(f"{path}: unbalanced {k} ({v:+d})")
            errs += 1
    return errs


if __name__ == "__main__":
    bad = sum(check(p) for p in sys.argv[1:])
    print("ok" if bad == 0 else f"{bad} problem(s)")
    sys.exit(1 if bad else 0)
