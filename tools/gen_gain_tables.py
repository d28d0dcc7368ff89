"""Prints the four gain tables the HT quantiser needs, 34 entries each, as C initialisers -- read from the reference's
QuantizerOJPH.cpp (T.814 / OpenJPH constants: BIBO gains of the 5/3 wavelet for the reversible exponents, square-root
energy gains of the 9/7 wavelet for the irreversible step sizes).  Run where /root/reference exists; the output is
pasted into grok_b200/csrc/geometry.cpp and oracle/j2k_oracle.c (constant tables, like tools/gen_ht_tables.py's)."""
import re, sys
src = open("/root/reference/src/lib/core/t2/quantizer/part15/QuantizerOJPH.cpp").read()
def table(cls, name):
    m = re.search(r"const float %s::%s\[34\] = \{(.*?)\};" % (cls, name), src, re.S)
    vals = [v.strip() for v in m.group(1).replace("\n", " ").split(",") if v.strip()]
    assert len(vals) == 34
    return vals
for out, cls, name in (("bibo_5x3_l", "bibo_gains", "gain_5x3_l"), ("bibo_5x3_h", "bibo_gains", "gain_5x3_h"),
                       ("sqe_9x7_l", "sqrt_energy_gains", "gain_9x7_l"), ("sqe_9x7_h", "sqrt_energy_gains", "gain_9x7_h")):
    v = table(cls, name)
    print("static const float %s[34] = {" % out)
    for i in range(0, 34, 7):
        print("    " + ", ".join(v[i:i + 7]) + ("," if i + 7 < 34 else ""))
    print("};")
