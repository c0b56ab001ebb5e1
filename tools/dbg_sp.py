import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bayeslogit_b200 import api
g = dict(np.load("tests/golden/pg_golden.npz"))
x, it = api.rpg_seeded("sp", g["psp_h"], g["psp_z"], 20240003)
bad = np.abs(x - g["psp_x"]) > 1e-12 * np.abs(g["psp_x"])
print("n", x.size, "bad", int(bad.sum()))
for i in np.nonzero(bad)[0]:
    print(i, "h", g["psp_h"][i], "z", g["psp_z"][i], "got", x[i], "want", g["psp_x"][i], "it", it[i], g["psp_iter"][i])
