#!/usr/bin/env python
"""Compare two builds of a translation unit kernel by kernel: floating-point opcode counts and (with --ops) the full multiset of
instructions with register numbers removed.   python tools/sass_ops_diff.py old.o new.o [--ops]"""
import collections, re, subprocess, sys


def funcs(obj):
    t = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    out, cur = {}, None
    for l in t.split("\n"):
        m = re.search(r"Function : (\S+)", l)
        if m:
            cur = m.group(1); out[cur] = []; continue
        if cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", l):
            out[cur].append(re.sub(r"/\* 0x[0-9a-f]+ \*/", "", l).strip())
    return out


def norm(body):
    o = []
    for l in body:
        l = re.sub(r"^/\*[0-9a-f]{4}\*/", "", l); l = re.sub(r"\.reuse", "", l)
        l = re.sub(r"\bU?R\d+\b", "R", l); l = re.sub(r"\bU?P\d\b", "P", l); l = re.sub(r"\bB\d+\b", "B", l)
        l = re.sub(r"0x[0-9a-f]+", "#", l)
        o.append(re.sub(r"\s+", " ", l).strip())
    return collections.Counter(o)


def opc(body):
    c = collections.Counter()
    for l in body:
        m = re.match(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", l)
        if m: c[m.group(1)] += 1
    return c


a, b = funcs(sys.argv[1]), funcs(sys.argv[2])
full = "--ops" in sys.argv
for k in sorted(set(a) | set(b)):
    name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()[:100]
    if k not in a or k not in b:
        print("ONLY-IN-%s %s" % ("OLD" if k in a else "NEW", name)); continue
    ca, cb = opc(a[k]), opc(b[k])
    keys = [o for o in sorted(set(ca) | set(cb)) if ca[o] != cb[o]]
    same_ops = norm(a[k]) == norm(b[k])
    print("%s %-100s n=%d/%d %s" % ("SAME " if same_ops else ("OPC= " if not keys else "DIFF "), name, len(a[k]), len(b[k]),
                                     {o: (ca[o], cb[o]) for o in keys} if keys else ""))
