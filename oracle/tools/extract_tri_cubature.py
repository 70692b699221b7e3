"""Extract the triangle cubature tables the reference reads at run time.

The reference (src/Basis2DTri.jl:274-296) loads `r s w` rows from
src/QuadratureData/quad_nodes_tri_N{N}.txt.  Those are numerical constants of
a published cubature rule; /root/reference does not exist on the GPU box, so
the numbers are stored once in oracle/data/tri_cubature.json.
Run in the build container only:  python oracle/tools/extract_tri_cubature.py
"""
import json, os, sys
import numpy as np

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/src/QuadratureData"
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "data", "tri_cubature.json")
tab = {}
for N in range(1, 17):  # degree 2N for N<=8 is all any config needs
    rsw = np.loadtxt(os.path.join(src, f"quad_nodes_tri_N{N}.txt"))
    tab[str(N)] = [[float(f"{v:.17g}") for v in row] for row in np.atleast_2d(rsw)]
with open(out, "w") as fh:
    json.dump(tab, fh)
print("wrote", os.path.normpath(out), {k: len(v) for k, v in tab.items()})
