#!/usr/bin/env python
"""Generate alloc_diag.inc / write_diag.inc for ref_driver.F90 from include/bgc_b200.h."""
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(HERE, "..", "..", "include", "bgc_b200.h")


def lists():
    src = open(HEADER).read()
    out = {}
    for m in re.finditer(r"#define\s+(\w+_LIST)\(X\)\s*\\\n((?:.*\\\n)*.*\n)", src):
        out[m.group(1)] = re.findall(r"X\((\w+)\)", m.group(2))
    return out


def main():
    L = lists()
    spec = [  # (list, fortran variable, dimension string)
        ("BGC_DIAG_K2_LIST", "bdiag", "(nL,nC)"), ("BGC_DIAG_KA_LIST", "bdiag", "(nL,nC,autotroph_cnt)"),
        ("BGC_DIAG_CA_LIST", "bdiag", "(nC,autotroph_cnt)"), ("BGC_DIAG_C1_LIST", "bdiag", "(nC)"),
        ("BGC_FLUX_DIAG_LIST", "bfdiag", "(nC)"), ("DMS_DIAG_LIST", "ddiag", "(nL,nC)"),
        ("DMS_FLUX_DIAG_LIST", "dfdiag", "(nC)"), ("MACROS_DIAG_LIST", "mdiag", "(nL,nC)")]
    with open(os.path.join(HERE, "alloc_diag.inc"), "w") as a, open(os.path.join(HERE, "write_diag.inc"), "w") as w:
        for lst, var, dim in spec:
            for name in L[lst]:
                a.write("  allocate(%s%%%s%s); %s%%%s = fill_value\n" % (var, name, dim, var, name))
                w.write("  write(ou) %s%%%s\n" % (var, name))
    print("wrote alloc_diag.inc, write_diag.inc")


if __name__ == "__main__":
    main()
