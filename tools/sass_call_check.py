#!/usr/bin/env python
"""sass_call_check.py -- static check of the out-of-line calls in the kernels' SASS (no GPU needed).

The streaming kernels call their exact-recompute path (cold_point) as a local subroutine; ptxas allocates registers across
such calls itself (no ABI).  On one instantiation it got that wrong: the callee chain overwrote a register that the calling
loop kept live across the call (the high word of the tile's byte offset) -> the stores after the call went to an illegal
address ("pressure-level tile loop").  This tool looks for that pattern in every kernel of an object file:

  for every CALL issued from the kernel body (not from inside a subroutine):
    clobber = registers written anywhere in the subroutine region and not reloaded from the stack there (LDL Rn, [R1+..])
    live    = registers READ before they are written on the straight-line path after the call's return
    report clobber & live

    python tools/sass_call_check.py earthkit-meteo_b200/csrc/build/lean/*.o

Straight-line scan from the return address, once around the enclosing loop; predicated writes do not kill: it can over-report
(a register that a skipped branch redefines), it does not miss a register that is read first on that path."""
import re
import subprocess
import sys

PAT = re.compile(r"\s+/\*([0-9a-f]{4,6})\*/\s+(@!?U?P[0-9T]\s+)?([A-Z0-9_.]+)\s*(.*?);")
NODEST = ("ST", "BRA", "BSSY", "BSYNC", "EXIT", "RET", "CALL", "NOP", "BAR", "MEMBAR", "WARPSYNC", "RED", "BREAK", "CCTL", "ERRBAR", "DEPBAR",
          "PREEXIT", "ACQBULK", "R2UR", "ISETP", "DSETP", "FSETP", "PLOP3", "R2P", "PREFETCH", "UISETP", "UMOV", "UIADD3", "ULOP3", "USHF", "UIMAD",
          "USEL", "ULDC", "LDCU", "S2UR", "ULEA", "UPLOP3", "UFLO", "UPRMT", "VOTEU", "UP2UR", "UR2UP", "HSETP2", "FCHK", "YIELD", "NANOSLEEP")


def regs(tok):
    return [int(x) for x in re.findall(r"\bR(\d+)\b", tok)]


def width(op):
    base = op.split(".")[0]
    if ".128" in op:
        return 4
    if ".64" in op or base in ("DFMA", "DMUL", "DADD", "DMNMX", "CS2R") or "WIDE" in op or op.startswith(("I2F.F64", "F2F.F64", "MUFU.RCP64H.X")):
        return 2
    return 1


def dest_src(op, operands):
    toks = [t.strip() for t in operands.split(",")] if operands else []
    base = op.split(".")[0]
    w = width(op)
    if any(base.startswith(n) for n in NODEST):
        src = [r for t in toks for r in regs(t)]
        if base.startswith("ST") and toks:
            d = regs(toks[-1])
            src += [d[-1] + k for k in range(w)] if d else []
        return [], src
    dst, src = [], []
    if len(toks) > 1 and re.fullmatch(r"U?P[0-9T]", toks[0]):  # "LOP3.LUT P2, R14, ..." : a predicate result first, then the register result
        toks = toks[1:]
    if toks and toks[0].startswith("R") and regs(toks[0]):
        dst = [regs(toks[0])[0] + k for k in range(w)]
        rest = toks[1:]
    else:
        rest = toks
    wide_src = base in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")
    for t in rest:
        for r in regs(t):
            src.append(r)
            if wide_src or ".64" in t:
                src.append(r + 1)
    return dst, src


def check(obj, pattern):
    names = [ln.split()[2] for ln in subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout.splitlines() if "Function :" in ln]
    n_calls = n_bad = 0
    for name in names:
        if pattern and not re.search(pattern, name):
            continue
        out = subprocess.run(["cuobjdump", "-sass", "-fun", name, obj], capture_output=True, text=True).stdout
        ins = []
        for ln in out.splitlines():
            m = PAT.match(ln)
            if m:
                ins.append((int(m.group(1), 16), (m.group(2) or "").strip(), m.group(3), m.group(4)))
        targets = sorted({int(m.group(1), 16) for a, p, op, opr in ins if op.startswith("CALL") for m in [re.search(r"0x([0-9a-f]+)", opr)] if m})
        if not targets:
            continue
        sub_lo = targets[0]  # subroutines are laid out after the kernel body, one region per call target
        bounds = targets + [ins[-1][0] + 16]
        region = {}  # target -> (registers written and not reloaded from the stack, targets called)
        for t0, t1 in zip(bounds[:-1], bounds[1:]):
            written, restored, callees = set(), set(), set()
            for a, p, op, opr in ins:
                if t0 <= a < t1:
                    d, _ = dest_src(op, opr)
                    written |= set(d)
                    if op.startswith("LDL") and re.search(r"\[R1(\+0x[0-9a-f]+)?\]", opr):
                        restored |= set(d)
                    if op.startswith("CALL"):
                        m = re.search(r"0x([0-9a-f]+)", opr)
                        if m:
                            callees.add(int(m.group(1), 16))
            region[t0] = (written - restored - {255}, callees)

        def clobber_of(t, seen_t=None):
            seen_t = seen_t or set()
            if t in seen_t or t not in region:
                return set()
            seen_t.add(t)
            w, cs = region[t]
            out_ = set(w)
            for c in cs:
                out_ |= clobber_of(c, seen_t)
            return out_

        index = {a: i for i, (a, _, _, _) in enumerate(ins)}
        # only the calls to cold_point-like subroutines (results through memory: they start by turning the generic pointers of their
        # array arguments into local-window offsets, c[0x0][0x2f8]); math slow paths called from the body return values in registers
        def takes_local_pointers(target):
            j = index.get(target)
            return j is not None and any("c[0x0][0x2f8]" in ins[k][3] for k in range(j, min(j + 8, len(ins))))

        for a, p, op, opr in ins:
            if not op.startswith("CALL") or a >= sub_lo:
                continue
            m = re.search(r"0x([0-9a-f]+)", opr)
            if not m or not takes_local_pointers(int(m.group(1), 16)):
                continue
            clobber = clobber_of(int(m.group(1), 16))
            n_calls += 1
            seen, hot, pred_written = set(), {}, {}
            i = index[a] + 1
            wrapped = False
            while i < len(ins) and ins[i][0] < sub_lo:
                a2, p2, op2, opr2 = ins[i]
                if wrapped and a2 >= a:
                    break  # once around the loop
                d, s = dest_src(op2, opr2)
                for r in s:
                    if r not in seen and p2 not in pred_written.get(r, ()):
                        seen.add(r)
                        if r in clobber:
                            hot.setdefault(r, f"{a2:#x} {op2} {opr2[:44]}")
                if not p2:
                    seen |= set(d)
                else:  # a predicated write defines the register for later reads under the same predicate
                    for r in d:
                        if r not in seen:
                            pred_written.setdefault(r, set()).add(p2)
                if op2 in ("EXIT",) and not p2:
                    break
                if op2.startswith("CALL"):  # a later call (e.g. the 64-bit division of the batched kernel) defines what its callee writes
                    m = re.search(r"0x([0-9a-f]+)", opr2)
                    if m:
                        seen |= clobber_of(int(m.group(1), 16))
                if op2 == "BRA":
                    m = re.search(r"0x([0-9a-f]+)", opr2)
                    tgt = int(m.group(1), 16) if m else None
                    if tgt is not None and tgt <= a and not wrapped:  # the back edge of the loop the call sits in: follow it once
                        wrapped = True
                        i = index.get(tgt, i + 1)
                        continue
                    if tgt is not None and not p2 and "UP" not in opr2 and tgt > a2:  # unconditional forward jump: follow it
                        if tgt not in index or (wrapped and tgt >= a):
                            break
                        i = index[tgt]
                        continue
                i += 1
            if hot:
                n_bad += 1
                print(f"{obj}: {name[:100]}\n   call at {a:#x}: live across the call AND written by the callee chain: " +
                      "; ".join(f"R{r} (read at {w})" for r, w in sorted(hot.items())))
    return n_calls, n_bad


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--only=")]
    only = next((a[7:] for a in sys.argv[1:] if a.startswith("--only=")), None)
    tot = bad = 0
    for o in args:
        c, b = check(o, only)
        tot += c
        bad += b
    print(f"{tot} call sites in kernel bodies checked, {bad} with a live register in the callee chain's write set")
