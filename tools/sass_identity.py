#!/usr/bin/env python
"""Are the default kernels of this tree instruction-for-instruction the ones of an earlier commit?

Used after adding the packed 2 x fp32 variants (template parameter PK, default false): the kernels that were measured and
parity-tested on B200 must not change.  Builds the three affected translation units of `ref` (default: the last commit whose build ran
on a B200) in a temporary directory, dumps the SASS of both builds with cuobjdump, drops the encodings, maps the old mangled names to
the new ones (the PK = false instantiations carry an extra `Lb0`), and compares the instruction streams.

  python tools/sass_identity.py [ref] > profiles/<round>_sass_identity.txt
"""
import os, re, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = sys.argv[1] if len(sys.argv) > 1 else "785f6fe"
UNITS = ("btkb_analysis", "btkb_synthesis", "btkb_perbin", "btkb_wide")
NVCC = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", "/usr/bin/g++", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-diag-suppress", "177"]


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


def ops(body):
    """Multiset of the instructions with register numbers, .reuse hints and branch targets removed."""
    out = []
    for l in body:
        l = re.sub(r"^/\*[0-9a-f]{4}\*/", "", l)
        l = re.sub(r"\.reuse", "", l)
        l = re.sub(r"\bU?R\d+\b", "R", l)
        l = re.sub(r"\bU?P\d\b", "P", l)
        l = re.sub(r"\bB\d+\b", "B", l)
        l = re.sub(r"0x[0-9a-f]+", "#", l)
        out.append(re.sub(r"\s+", " ", l))
    return sorted(out)


def new_name(n):
    if "k_perbinI" in n:
        n = re.sub(r"(k_perbinILi\d+ELi\d+ELi\d+)E", r"\1ELb0E", n)
    n = re.sub(r"(k_covarianceILi\d+)E", r"\1ELb0E", n)
    n = re.sub(r"(k_perbin_rlsILi\d+)E", r"\1ELb0E", n)
    n = re.sub(r"(k_perbin_wideILi\d+ELi\d+)E", r"\1ELb0E", n)
    if "k_analysis_r1" in n:
        n = n.replace("EEEvNS_12AnalysisArgsE", "ELb0EEEvNS_12AnalysisArgsE")
    if "k_synthesis_fast" in n:
        n = n.replace("EEEvNS_13SynthesisArgsE", "ELb0EEEvNS_13SynthesisArgsE")
    return n


def main():
    total = same = ralloc = 0
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run("git -C %s archive %s distant_speech_recognition_b200/csrc include | tar -x -C %s" % (ROOT, REF, tmp), shell=True, check=True)
        old_src = os.path.join(tmp, "distant_speech_recognition_b200", "csrc")
        new_src = os.path.join(ROOT, "distant_speech_recognition_b200", "csrc")
        print("reference commit %s vs working tree (%s)" % (REF, subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()))
        for u in UNITS:
            subprocess.run(NVCC + ["-c", u + ".cu", "-o", os.path.join(tmp, u + "_old.o")], cwd=old_src, check=True)
            subprocess.run(NVCC + ["-c", u + ".cu", "-o", os.path.join(tmp, u + "_new.o")], cwd=new_src, check=True)
            old, new = funcs(os.path.join(tmp, u + "_old.o")), funcs(os.path.join(tmp, u + "_new.o"))
            for name, body in sorted(old.items()):
                nb = new.get(new_name(name))
                ok = nb == body
                total += 1; same += ok
                tag = "same" if ok else "DIFF"
                if not ok and nb is not None and ops(nb) == ops(body):
                    tag = "ralloc"; ralloc += 1   # same multiset of operations once register numbers and branch targets are dropped: only allocation / order moved
                print("%s  %-6s %5d instructions  %s" % (u, tag, len(body), name))
            extra = sorted(set(new) - {new_name(n) for n in old})
            for name in extra:
                print("%s  new   %5d instructions  %s" % (u, len(new[name]), name))
    print("default kernels identical: %d of %d (+ %d with the same operations under a different register allocation)" % (same, total, ralloc))
    return 0 if same + ralloc == total else 1


if __name__ == "__main__":
    sys.exit(main())
