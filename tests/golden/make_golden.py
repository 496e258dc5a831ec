"""Generates the golden fixtures under tests/golden/ (run in the build container,
where /root/reference exists; the GPU box never needs /root/reference).

  lut_ref.npz : the two 1024-entry tables parsed from the reference's own
                src/dotp_lut.h, and the same tables produced by the reference's
                generator src/mk_lut.cpp compiled by oracle/build_ref.sh.
"""
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"


def parse_lut(path=None, txt=None):
    txt = open(path).read() if txt is None else txt
    out = {}
    for name in ("dotp_lut_a", "dotp_lut_b"):
        m = re.search(name + r"\[1024\][^=]*=\s*\{(.*?)\};", txt, re.S)
        vals = [float(v) for v in re.findall(r"^\s*([0-9.]+)\s*,", m.group(1), re.M)]
        assert len(vals) == 1024, (name, len(vals))
        out[name] = np.array(vals)
    return out


def main():
    ref = parse_lut(os.path.join(REF, "src/dotp_lut.h"))
    import subprocess
    gen = parse_lut(txt=subprocess.run([os.path.join(ROOT, "oracle/_ref/mk_lut")], capture_output=True, text=True, check=True).stdout)
    for k in ref:
        assert np.array_equal(ref[k], gen[k]), k
    np.savez_compressed(os.path.join(ROOT, "tests/golden/lut_ref.npz"), a=ref["dotp_lut_a"].astype(np.int8), b=ref["dotp_lut_b"].astype(np.int8))
    print("lut_ref.npz written; dotp_lut.h == mk_lut output")


if __name__ == "__main__":
    sys.exit(main())
