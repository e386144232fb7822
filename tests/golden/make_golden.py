"""Regenerates the golden fixtures from the reference tree (run in the build container only;
/root/reference does not exist on the GPU box, the committed .npz files travel instead).

  espp_positions.npz  <- tests/LennardJones/positions.txt  (32 768 x 3 doubles, reader semantics of
                         mrmd/io/RestoreTXT.cpp:24-74) + the constants of
                         tests/LennardJones/LennardJones.cpp:32-38,88,96,108
  lj_nvt_final.npz    <- examples/02_LennardJones_NVE/lennardJonesNVT_final.gro (4096 atoms, pos+vel, box;
                         reader semantics of mrmd/io/RestoreGRO.cpp:24-147: fixed-width
                         "%5d%5s%5s%5d%8lf%8lf%8lf%8lf%8lf%8lf", last line = box)
"""
import os
import sys

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    pos = np.loadtxt(os.path.join(REF, "tests/LennardJones/positions.txt"), dtype=np.float64)
    assert pos.shape == (32768, 3)
    np.savez_compressed(
        os.path.join(HERE, "espp_positions.npz"), pos=pos, box=np.array([33.8585] * 3), rc=2.5, skin=0.3,
        espp_real=32768, espp_ghost=22104, espp_neighbors=1310403, nonperiodic_pairs_with_ghosts=1426948,
        espp_initial_energy=-94795.927)

    with open(os.path.join(REF, "examples/02_LennardJones_NVE/lennardJonesNVT_final.gro")) as f:
        lines = f.read().split("\n")
    n = int(lines[1])
    p = np.zeros((n, 3))
    v = np.zeros((n, 3))
    for i in range(n):
        ln = lines[2 + i]
        vals = [float(ln[20 + 8 * k: 28 + 8 * k]) for k in range(6)]
        p[i] = vals[:3]
        v[i] = vals[3:]
    box = np.array([float(t) for t in lines[2 + n].split()])
    np.savez_compressed(os.path.join(HERE, "lj_nvt_final.npz"), pos=p, vel=v, box=box)
    print("wrote fixtures:", n, "gro atoms;", pos.shape[0], "espp atoms")


if __name__ == "__main__":
    main()
