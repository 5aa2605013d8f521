#!/usr/bin/env python
"""Generate Nyquist(M) analysis/synthesis prototypes with the REFERENCE's own design tool.

Runs /root/reference/btk20_src/tools/filterbank/design_nyquist_filter.py in-process (this container only;
/root/reference does not exist on the GPU box) after aliasing the NumPy-1 names the tool still uses
(`np.float_`, `np.float`) to `np.float64` — the two-token patch SURVEY.md §8c verified reproduces the shipped
M=256 pickles to 4e-13.  Output: tests/golden/prototype_M{M}_m{m}_r{r}.npz with float64 arrays h, g.

The M=256 prototypes are additionally copied from the reference's shipped pickles
(unit_test/prototype.ny/{h,g}-M256-m4-r1.pickle) so the test-suite can check the tool against them.

Usage: python tests/golden/make_prototypes.py [M ...]      (default: 256 512 1024)
"""
import os, sys, pickle, importlib.util, tempfile
import numpy as np

REF = "/root/reference/btk20_src"
HERE = os.path.dirname(os.path.abspath(__file__))


def load_tool():
    if not hasattr(np, "float_"):
        np.float_ = np.float64
    if not hasattr(np, "float"):
        np.float = np.float64
    spec = importlib.util.spec_from_file_location("design_nyquist_filter", os.path.join(REF, "tools/filterbank/design_nyquist_filter.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main(Ms, m=4, r=1):
    tool = load_tool()
    for M in Ms:
        with tempfile.TemporaryDirectory() as td:
            tool.main(M, m, r, td)
            h = np.asarray(pickle.load(open(os.path.join(td, "h-M%d-m%d-r%d.pickle" % (M, m, r)), "rb")), np.float64)
            g = np.asarray(pickle.load(open(os.path.join(td, "g-M%d-m%d-r%d.pickle" % (M, m, r)), "rb")), np.float64)
        out = os.path.join(HERE, "prototype_M%d_m%d_r%d.npz" % (M, m, r))
        np.savez(out, h=h, g=g)
        print("wrote", out, h.shape, g.shape)
    # the shipped fixtures (python-2 pickles)
    hs = np.asarray(pickle.load(open(os.path.join(REF, "unit_test/prototype.ny/h-M256-m4-r1.pickle"), "rb"), encoding="latin1"), np.float64)
    gs = np.asarray(pickle.load(open(os.path.join(REF, "unit_test/prototype.ny/g-M256-m4-r1.pickle"), "rb"), encoding="latin1"), np.float64)
    np.savez(os.path.join(HERE, "prototype_shipped_M256_m4_r1.npz"), h=hs, g=gs)


if __name__ == "__main__":
    Ms = [int(a) for a in sys.argv[1:]] or [256, 512, 1024]
    main(Ms)
