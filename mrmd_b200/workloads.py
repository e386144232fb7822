"""Synthetic lattice-initialised systems of BASELINE.json / SURVEY.md section 8(d) (host-side input
construction only: plain numpy, no compute path)."""
import numpy as np

RHO_SPACING = 1.25  # simple-cubic spacing for rho = 0.512


def lattice_system(n_side, spacing=RHO_SPACING, seed=1234, max_velocity=1.0, n_side_x=None):
    """Simple-cubic lattice, positions (i + 1/2) * spacing, velocities (u - 1/2) * max_velocity from a
    Philox stream keyed by `seed`, net momentum removed.  n_side_x != n_side gives an x-elongated box (weak
    scaling over x-slabs)."""
    nx = n_side_x or n_side
    idx = np.stack(np.meshgrid(np.arange(nx), np.arange(n_side), np.arange(n_side), indexing="ij"), axis=-1)
    pos = (idx.reshape(-1, 3) + 0.5) * spacing
    rng = np.random.Generator(np.random.Philox(key=seed))
    vel = (rng.random(pos.shape) - 0.5) * max_velocity
    vel -= vel.mean(axis=0, keepdims=True)
    box = np.array([nx * spacing, n_side * spacing, n_side * spacing])
    return pos, vel, box


CONFIGS = {
    # name: (n_side, description)
    "lj_nvt_1m": (100, "Lennard-Jones NVT (examples/01 physics, examples/02 rebuild loop) scaled to 1M atoms: "
                       "sc lattice 100^3, L=125, rho=0.512, rc=2.5, skin=0.1, r_cap=0.7, Langevin gamma=20 T=1.5, "
                       "dt=0.002"),
    "lj_nvt_4k": (16, "examples/02 size: 4096 atoms, L=20"),
    "adress_8m": (200, "LJ / ideal-gas AdResS slab scaled to 8M atoms with thermodynamic force"),
}

TETRAMER_SPACING = 1.98425  # molecule density 0.128 -> atom density 0.512


def tetramer_system(n_side, spacing=TETRAMER_SPACING, seed=1234, max_velocity=1.0, n_side_x=None):
    """n_side^3 tetramers: centres of mass on a simple-cubic lattice, each a regular tetrahedron of edge 1 centred on its
    site, atoms of a molecule contiguous (atomsOffset = 4 m); one velocity per molecule (no velocity along the bonds),
    net momentum removed.  Returns pos[4M,3], vel[4M,3], box."""
    nx = n_side_x or n_side
    idx = np.stack(np.meshgrid(np.arange(nx), np.arange(n_side), np.arange(n_side), indexing="ij"), axis=-1).reshape(-1, 3)
    sites = (idx + 0.5) * spacing
    tet = np.array([(1, 1, 1), (1, -1, -1), (-1, 1, -1), (-1, -1, 1)], dtype=np.float64) / (2.0 * np.sqrt(2.0))
    pos = (sites[:, None, :] + tet[None, :, :]).reshape(-1, 3)
    rng = np.random.Generator(np.random.Philox(key=seed))
    vmol = (rng.random(sites.shape) - 0.5) * max_velocity
    vmol -= vmol.mean(axis=0, keepdims=True)
    vel = np.repeat(vmol, 4, axis=0)
    return pos, vel, np.array([nx * spacing, n_side * spacing, n_side * spacing])
