"""Generate the committed fixtures under tests/golden/ from the reference tree.

Run in the build container (where /root/reference exists):
    python tests/golden/make_golden.py
Inputs are the reference's own shipped data files; outputs are
  arm_data.npz      raw train (10 trials) / val (5 trials) of the 3-link arm (configs 1, 2)
  snake_data.npz    raw train / val of snake-data.mat (config 3)
  rsys_subset.npz   first 3 systems of rand-systems_2021-01-10_16-59 (config 4 parity cases)
  arm_blockM_Z.npz  the reference's OWN lifted states res_lin.Z / res_bilin.Z with res_*.Y
                    (golden vectors for scaling + poly dictionary + PCA + econ layout)
Nothing here is produced by the oracle: these are reference inputs and reference outputs.
"""
import os
import sys

import numpy as np
import scipy.io as sio

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle as O  # noqa: E402  (loaders only)

REF = "/root/reference/"


def pack(d, prefix=""):
    out = {}
    for split in ("train", "val"):
        out[f"{prefix}{split}_n"] = np.array(len(d[split]))
        for i, tr in enumerate(d[split]):
            for k in ("t", "y", "u"):
                out[f"{prefix}{split}{i}_{k}"] = tr[k]
    return out


arm = O.load_data4sysid(REF + "datafiles/arm-3link-markers-noload-50trials_train-10_val-5.mat")
np.savez_compressed(os.path.join(HERE, "arm_data.npz"), **pack(arm))
snake = O.load_data4sysid(REF + "datafiles/snake-data.mat")
np.savez_compressed(os.path.join(HERE, "snake_data.npz"), **pack(snake))
rs = O.load_rand_systems(REF + "datafiles/rand-systems_2021-01-10_16-59 (1)/rsys-all_train-9_val-1.mat")
out = {"nsys": np.array(3)}
for i in range(3):
    out.update(pack(rs[i], prefix=f"s{i}_"))
np.savez_compressed(os.path.join(HERE, "rsys_subset.npz"), **out)
sim = REF + "systems/thesis-arm-markers_noload_3-mods_1-links_20hz/simulations/blockM_c0p45-0p35_0p5x0p5_15sec/"
lin = sio.loadmat(sim + "linear_poly-3_n-6_m-3_del-0_2020-06-09_16-42.mat", squeeze_me=True, struct_as_record=False)["res_lin"]
bil = sio.loadmat(sim + "bilinear_poly-3_n-6_m-3_del-0_2020-06-09_16-43.mat", squeeze_me=True, struct_as_record=False)["res_bilin"]
np.savez_compressed(os.path.join(HERE, "arm_blockM_Z.npz"), lin_Y=lin.Y, lin_Z=lin.Z, bil_Y=bil.Y, bil_Z=bil.Z)
for f in sorted(os.listdir(HERE)):
    print(f, os.path.getsize(os.path.join(HERE, f)))
