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

# Config 4 at file level: one random system from EVERY shipped `rsys-all` file (35 files, 292 systems in all) plus
# two of the 20 single-system files of rand-systems_2021-01-11_11-24 (50 training trials each).  The full 312-system set is
# 56 MB of float64 trajectories — too large to commit — so the fixture samples every FILE instead of every system.
import glob  # noqa: E402

out = {}
k = 0
src = []
for d in sorted(glob.glob(REF + "datafiles/rand-systems_*")):
    f = sorted(glob.glob(d + "/rsys-all*.mat"))
    if not f:
        continue
    rs = O.load_rand_systems(f[0])
    i = len(rs) - 1                                   # the LAST system of the file (rsys_subset.npz holds the first three of file 1)
    out.update(pack(rs[i], prefix=f"s{k}_"))
    src.append(f"{os.path.basename(d)}/{os.path.basename(f[0])}#{i}")
    k += 1
single = REF + "datafiles/rand-systems_2021-01-11_11-24 (1)/"
for name in ("rsys-1_train-50_val-1.mat", "rsys-20_train-50_val-1.mat"):
    out.update(pack(O.load_data4sysid(single + name), prefix=f"s{k}_"))       # a single-system file is a plain data4sysid
    src.append(f"rand-systems_2021-01-11_11-24 (1)/{name}#0")
    k += 1
out["nsys"] = np.array(k)
out["source"] = np.array(src)
np.savez_compressed(os.path.join(HERE, "rsys_files.npz"), **out)
print("rsys_files.npz", k, "systems", os.path.getsize(os.path.join(HERE, "rsys_files.npz")))
