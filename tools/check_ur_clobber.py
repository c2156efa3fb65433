#!/usr/bin/env python3
"""Scan the SASS of liblokib200.so for a ptxas (12.9) register-allocation fault seen once in this code base: a warp reduction
(CREDUX / REDUX, i.e. __reduce_*_sync) writes its result into a uniform register that still holds a live value -- in the observed case
the shared-memory base used by the next LDS -- so the load goes to a wild address (compute-sanitizer: "Invalid __shared__ read ...
misaligned" in k_advance_stream<F_ECR>).  The pattern checked: after `CREDUX/REDUX URn`, URn is used as an ADDRESS operand
([R+URn+imm], or the base of a LEA that feeds a shared access) before anything writes URn again.
usage: python tools/check_ur_clobber.py [lib.so]  -> exit status 1 when a suspect is found"""
import re
import subprocess
import sys


_SASS = {}


def sass(path):
    if path not in _SASS:
        _SASS[path] = subprocess.run(["cuobjdump", "-sass", path], stdout=subprocess.PIPE, text=True).stdout
    return _SASS[path]


def scan(path):
    out = sass(path)
    func, window, bad = None, {}, []
    ins = re.compile(r"/\*([0-9a-f]{4,})\*/\s+(.*?);")
    for line in out.split("\n"):
        if "Function :" in line:
            func, window = line.split("Function :")[1].strip(), {}
            continue
        m = ins.search(line)
        if not m:
            continue
        addr, text = m.group(1), m.group(2).strip()
        ops = text.split()
        opcode = ops[1] if ops and ops[0].startswith("@") else (ops[0] if ops else "")
        dst = re.search(r"\b(UR\d+)\b", text)
        if opcode.startswith(("CREDUX", "REDUX")) and dst:
            window[dst.group(1)] = (addr, 0)
            continue
        for ur in list(window):
            start, age = window[ur]
            used_as_address = re.search(r"\[[^\]]*\b%s\b[^\]]*\]" % ur, text) or (opcode.startswith(("LEA", "ULEA")) and re.search(r",\s*%s\b" % ur, text))
            writes = re.match(r"(@!?U?P\d+\s+)?\S+\s+%s\b" % ur, text) is not None and not opcode.startswith(("ST", "ATOM", "RED"))
            if used_as_address:
                bad.append((func, start, addr, ur, text))
                del window[ur]
            elif writes or age > 40 or (not ops[0].startswith("@") and opcode in ("BRA", "EXIT", "RET", "CALL")):
                del window[ur]
            else:
                window[ur] = (start, age + 1)
    return bad


def scan_derived_bases(path, kernel="k_advance_stream"):
    """Second fault seen with the same ptxas: a uniform register holding `shared base + 16 P` (an array whose offset depended on the
    process count) was later used as the plain base.  The streaming kernel now keeps every shared array at a compile-time offset, so
    its SASS must contain no uniform LEA other than the base computation itself (ULEA URx, URcga, URx, 0x18)."""
    out = sass(path)
    func, bad = None, []
    for line in out.split("\n"):
        if "Function :" in line:
            func = line.split("Function :")[1].strip()
            continue
        if func and kernel in func and "ULEA" in line:
            m = re.search(r"ULEA\S*\s+(UR\d+),\s*(UR\d+),\s*(UR\d+),\s*(0x[0-9a-f]+)", line)
            if m and not (m.group(1) == m.group(3) and m.group(4) == "0x18"):
                bad.append((func, line.strip()[:100]))
    return bad


if __name__ == "__main__":
    import os
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "loki_mc_b200", "liblokib200.so")
    bad = scan(lib)
    for f, a0, a1, ur, text in bad:
        print("SUSPECT %s: reduction at %s writes %s, used as an address at %s: %s" % (f, a0, ur, a1, text))
    derived = scan_derived_bases(lib)
    for f, text in derived:
        print("DERIVED BASE %s: %s" % (f, text))
    print("%d suspect site(s), %d derived shared bases in the streaming kernel" % (len(bad), len(derived)))
    sys.exit(1 if bad or derived else 0)
