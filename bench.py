#!/usr/bin/env python
"""bench.py -- atom-steps/s (and pair-interactions/s) of the MRMD force + neighbour hot path on B200.

  python bench.py --gpus N --steps K --warmup W            our arm (hand-written sm_100a kernels via the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  the reference's CPU path (OpenMP restatement: the
                                                           reference itself needs Kokkos/Cabana, unbuildable offline)

A "step" is one MD time step of the workload (pre-force integrate, ghost refresh or neighbour rebuild, force
zero, LJ force, ghost fold-back, post-force integrate).  One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "atom-steps/s"
PHYS = dict(dt=0.002, rc=2.5, skin=0.1, sigma=1.0, epsilon=1.0, cap=0.7, max_neigh=60, zeta=20.0, temperature=1.5,
            seed=1234)


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=2000)
    p.add_argument("--warmup", type=int, default=300)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--side", type=int, default=100, help="lattice sites per box edge per GPU (100 -> 1M atoms)")
    p.add_argument("--side-x", type=int, default=None, help="lattice sites per GPU along x if different from --side "
                   "(configs[4]: --workload adress --side 400 --side-x 50 --gpus 8 = 64M atoms in a 500^3 box)")
    p.add_argument("--equil", type=int, default=300, help="untimed equilibration steps that melt the lattice")
    p.add_argument("--full-list", type=int, default=2, help="0 half list, 1 full list, 2 tiled periodic full list (fast path)")
    p.add_argument("--e2e-steps", type=int, default=100)
    p.add_argument("--cpu-seconds", type=float, default=20.0, help="budget of the cpu_baseline leg")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--replicas", action="store_true", help="N>1: independent periodic replicas instead of x-slabs")
    p.add_argument("--balance", action="store_true", help="adress workload on N>1 GPUs: cost-balanced slab widths")
    p.add_argument("--force-cost-ratio", type=float, default=4.5,
                   help="--balance: force-kernel time / rest of the step, per atom of the AT + HY region")
    p.add_argument("--workload", default="lj", choices=["lj", "adress", "tetramer"],
                   help="lj: configs[1] (the headline line); adress: configs[2]/[4] physics (LJ / ideal-gas AdResS slab "
                        "with thermodynamic force, one molecule per atom; use --side 200 for 8M atoms per GPU); tetramer: "
                        "configs[3] physics (AdResS LJ tetramers, spherical region, SHAKE / RATTLE; --side = molecules "
                        "per edge, 160 -> 16.4M atoms; one GPU, or --replicas)")
    args = p.parse_args()
    if args.workload == "tetramer":
        args.full_list = 0  # half Verlet list of molecules over materialised ghost molecules (reference semantics)
        if args.gpus > 1 and not args.replicas:
            p.error("--workload tetramer: the x-slab decomposition handles one-atom molecules only; use --replicas")
    return args


def env_rank():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    """dram bytes per launch of the force kernel from the committed ncu capture, if any."""
    path = os.path.join(ROOT, "profiles", "force_kernel_traffic.json")
    if os.path.exists(path):
        try:
            return json.load(open(path))
        except Exception:
            pass
    return None


def cpu_loop(pos, vel, box, steps, warmup, threads):
    """The reference path's OpenMP restatement on the host cores (oracle/ is only ever the baseline/checker)."""
    from oracle import pyoracle as orc
    from oracle.md_loop import OracleMD

    orc.build()
    orc.lib().or_set_threads(threads)
    md = OracleMD(pos, vel, box, dt=PHYS["dt"], rc=PHYS["rc"], skin=PHYS["skin"], sigma=PHYS["sigma"],
                  epsilon=PHYS["epsilon"], cap=PHYS["cap"], max_neigh=PHYS["max_neigh"], langevin=True,
                  zeta=PHYS["zeta"], temperature=PHYS["temperature"], seed=PHYS["seed"], cell_sort=True)
    md.run(warmup)
    return md.run(steps), md


ADRESS_THERMO = dict(targetDensity=0.512, binWidth=0.25, modulation=2.0, sampleInterval=10, updateInterval=1000,
                     sigma=2.0, range=2.0)


def cpu_loop_adress(pos, vel, box, steps, warmup, threads):
    """--workload adress on the host cores: the section 3.5 AdResS step through the oracle (checker / baseline only)."""
    from oracle import pyoracle as orc
    from oracle.md_loop import OracleAdressMD

    orc.build()
    orc.lib().or_set_threads(threads)
    lx = float(box[0])
    weight = orc.make_weight(orc.WEIGHT_SLAB, [lx / 2, box[1] / 2, box[2] / 2], 0.2 * lx, 0.1 * lx, 1)
    md = OracleAdressMD(pos, vel, box, weight, dt=PHYS["dt"], rc=PHYS["rc"], skin=PHYS["skin"], sigma=PHYS["sigma"],
                        epsilon=PHYS["epsilon"], cap=PHYS["cap"], max_neigh=PHYS["max_neigh"], langevin=True,
                        zeta=PHYS["zeta"], temperature=PHYS["temperature"], seed=PHYS["seed"], thermo=ADRESS_THERMO)
    md.run(warmup)
    return md.run(steps), md


TETRAMER = dict(constraint_iterations=3, bond_length=1.0, max_neigh=40)


def tetramer_region(box):
    """config 4: spherical region, centre = box centre, R = 60 and h = 30 in the 317.48 box (scaled with the box)"""
    return 60.0 / 317.48 * float(box[0]), 30.0 / 317.48 * float(box[0])


def cpu_loop_tetramer(pos, vel, box, steps, warmup, threads):
    """--workload tetramer on the host cores (checker / baseline only)."""
    from oracle import pyoracle as orc
    from oracle.md_loop import OracleAdressMD

    orc.build()
    orc.lib().or_set_threads(threads)
    radius, hybrid = tetramer_region(box)
    weight = orc.make_weight(orc.WEIGHT_SPHERICAL, 0.5 * np.asarray(box), radius, hybrid, 2)
    md = OracleAdressMD(pos, vel, box, weight, dt=PHYS["dt"], rc=PHYS["rc"], skin=PHYS["skin"], sigma=PHYS["sigma"],
                        epsilon=PHYS["epsilon"], cap=PHYS["cap"], max_neigh=TETRAMER["max_neigh"], langevin=True,
                        zeta=PHYS["zeta"], temperature=PHYS["temperature"], seed=PHYS["seed"], atoms_per_mol=4,
                        constraint_iterations=TETRAMER["constraint_iterations"], bond_length=TETRAMER["bond_length"])
    md.run(warmup)
    return md.run(steps), md


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    """--impl reference: rank 0 only; a bounded sample (smaller periodic system at the same state point) so that
    --steps K --warmup W ends within a few minutes; atom-steps/s is size normalised."""
    rank, _, world = env_rank()
    if rank != 0:
        return
    from mrmd_b200.workloads import lattice_system

    threads = host_threads()
    # probe the host rate on a 32^3 system, then size the sample for ~150 s of CPU work
    loop = {"lj": cpu_loop, "adress": cpu_loop_adress, "tetramer": cpu_loop_tetramer}[args.workload]
    if args.workload == "tetramer":
        from mrmd_b200.workloads import tetramer_system

        def lattice_system(side):  # noqa: F811  (molecules per edge; same probe / sizing logic in atoms)
            return tetramer_system(side)
    pos, vel, box = lattice_system(32 if args.workload != "tetramer" else 20)
    probe, _ = loop(pos, vel, box, 6, 2, threads)
    rate = len(pos) * probe["steps"] / probe["seconds"]
    budget_atoms = rate * 150.0 / max(args.steps + args.warmup, 1)
    if args.workload == "tetramer":
        budget_atoms /= 4.0
    side = int(max(16, min(args.side, np.floor(budget_atoms ** (1.0 / 3.0)))))
    pos, vel, box = lattice_system(side)
    n = len(pos)
    res, md = loop(pos, vel, box, args.steps, args.warmup, threads)
    value = n * res["steps"] / res["seconds"]
    sample = (f"periodic sc-lattice system of {n} atoms ({side}^3, same rho/T/dt/skin as the {args.side}^3 workload), "
              f"{args.warmup} warm-up + {args.steps} timed steps, {res['rebuilds']} neighbour rebuilds")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "atom-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * res["seconds"] / max(res["steps"], 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, n_atoms_per_gpu=args.side ** 3 * (4 if args.workload == "tetramer" else 1)),
        "pair_interactions_per_s": res["pairInteractions"] / res["seconds"],
        "cpu_baseline": {"value": value, "unit": "atom-steps/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "atom-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "OpenMP restatement of the reference path (Kokkos 4.7.1 / Cabana 0.7 un-vendored: the reference "
                "cannot be built offline)",
    }
    emit(line)


def workload_config(args, n_atoms_per_gpu):
    lj = ("Lennard-Jones NVT (examples/01 physics, examples/02 rebuild loop, tests/NVT spatial sort) scaled "
          "to 1M atoms per GPU: sc lattice, rho=0.512, rc=2.5 sigma, skin 0.1, r_cap 0.7, Langevin gamma=20 "
          "T=1.5, dt=0.002, maxNeighbors 60")
    ad = ("LJ / ideal-gas AdResS slab (examples/04 physics + the section 3.5 AdResS step): one molecule per atom, sc "
          "lattice rho=0.512, Slab(centre, AT 0.2 Lx, HY 0.1 Lx, nu 1), LJ_IdealGas(cap 0.7, rc 2.5, shift), "
          "ThermodynamicForce(rho 0.512, bin 0.25, modulation 2, sample/10, update/1000), Langevin gamma=20 T=1.5, "
          "dt=0.002, skin 0.1, maxNeighbors 60")
    tet = ("AdResS Lennard-Jones tetramers (configs[3]): molecule centres on an sc lattice of spacing 1.98425 (atom density "
           "0.512), regular tetrahedra of edge 1, 4 atoms per molecule (relMass 1/4), Spherical(centre, R 0.189 L, h 0.0945 "
           "L, exponent 2) = R 60 / h 30 at 16.4M atoms, LJ_IdealGas(cap 0.7, rc 2.5, shift), MoleculeConstraints(4, 3) "
           "on the six bonds (SHAKE / RATTLE), Langevin gamma=20 T=1.5, dt=0.002, skin 0.1, half Verlet list of molecules "
           "over MultiResGhostLayer ghosts, maxNeighbors 40")
    wl = getattr(args, "workload", "lj")
    sp = 1.98425 if wl == "tetramer" else 1.25
    return {
        "workload": {"lj": lj, "adress": ad, "tetramer": tet}[wl],
        "atoms_per_gpu": n_atoms_per_gpu, "box_per_gpu": [(getattr(args, "side_x", None) or args.side) * sp, args.side * sp, args.side * sp], "equilibration_steps": args.equil,
        "list": {0: "half (reference semantics, fp64 RED scatter)", 1: "full (generic gather kernel)", 2: "full, periodic tiles staged in shared memory (mrmd_b200_verlet_build_periodic)"}[args.full_list],
        "l2": "inputs larger than L2 (per step: 104 B/atom state + neighbour table ~ 4 B x 19-38 slots/atom > 126 MB "
              "at 1M atoms); no explicit flush",
        "parallelism": "1 GPU" if args.gpus == 1 else (
            f"{args.gpus} independent periodic replicas, one per GPU" if getattr(args, "replicas", False) else
            f"{args.gpus} x-slabs of one {args.gpus * (getattr(args, 'side_x', None) or args.side) * 1.25:g} x {args.side * 1.25:g} x {args.side * 1.25:g} box, "
            "one process per GPU, NCCL halos (positions every step, full records at rebuild), ncclAllReduce(max) "
            "rebuild decision"),
    }


def run_b200(args):
    import torch

    rank, local_rank, world = env_rank()
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from mrmd_b200 import api
    from mrmd_b200.workloads import lattice_system

    api.L().mrmd_b200_set_device(local_rank)
    stream = torch.cuda.current_stream().cuda_stream

    slab_mode = world > 1 and not args.replicas
    tetramer = args.workload == "tetramer"
    adress = args.workload in ("adress", "tetramer")
    sites_x = args.side_x or args.side          # lattice planes per GPU along x (equal-width slabs)
    spacing = 1.25
    global_lx = world * sites_x * spacing if slab_mode else sites_x * spacing
    cuts = None
    my_sites_x, x_offset = sites_x, (rank * sites_x * spacing if slab_mode else 0.0)
    if slab_mode and adress and args.balance:
        # cost-balanced slabs: the AT + HY region [0.3 Lx, 0.7 Lx] carries the force kernel, the
        # coarse-grained rest only streams; boundaries on lattice planes so that every rank builds its own atoms
        from mrmd_b200 import slabs

        cuts = slabs.balanced_cuts([0.0], [global_lx], world, 0.3 * global_lx, 0.7 * global_lx, args.force_cost_ratio,
                                   quantum=spacing, min_width=2 * (PHYS["rc"] + PHYS["skin"]))
        my_sites_x = int(round((cuts[rank + 1] - cuts[rank]) / spacing))
        x_offset = float(cuts[rank])
    if tetramer:
        from mrmd_b200.workloads import tetramer_system

        pos, vel, box = tetramer_system(args.side, seed=PHYS["seed"] + rank)
    else:
        pos, vel, box = lattice_system(args.side, seed=PHYS["seed"] + rank, n_side_x=my_sites_x)
        box = np.array([sites_x * spacing, box[1], box[2]])  # the equal-width slab: world * box[0] is the global length
    n = len(pos)
    sub = api.Subdomain([0, 0, 0], box, PHYS["rc"] + PHYS["skin"])
    if slab_mode:
        # weak scaling over x-slabs: one global box of world * sites_x x side x side sites, rank r owns slab r
        pos = pos + np.array([x_offset, 0.0, 0.0])
    atoms = api.Atoms.from_arrays(pos, vel, mass=1.0, relativeMass=0.25 if tetramer else 1.0)

    extra = {}
    if tetramer:
        radius, hybrid = tetramer_region(box)
        extra = dict(adress=True, weight=api.Spherical(0.5 * box, radius, hybrid, 2), doShift=True, atomsPerMolecule=4,
                     numConstraintIterations=TETRAMER["constraint_iterations"], bondLength=TETRAMER["bond_length"])
    elif adress:
        # configs[2] / configs[4] physics: Slab(centre of the global box, AT diameter 0.2 Lx, HY width 0.1 Lx, nu = 1),
        # LJ_IdealGas(cap 0.7, rc 2.5, shift on), ThermodynamicForce(rho 0.512, bin 0.25, modulation 2; sample every
        # 10 steps, update(sigma 2, range 2) every 1000 steps), one molecule per atom
        lx = box[0] * (world if slab_mode else 1)
        extra = dict(adress=True, weight=api.Slab([lx / 2, box[1] / 2, box[2] / 2], 0.2 * lx, 0.1 * lx, 1), doShift=True,
                     thermo=dict(targetDensity=0.512, binWidth=0.25, modulation=2.0, sampleInterval=10,
                                 updateInterval=1000, sigma=2.0, range=2.0))

    def make_md(a):
        if slab_mode:
            from mrmd_b200 import slabs

            uid = slabs.broadcast_unique_id(rank)
            return slabs.SlabMolecularDynamics(a, np.zeros(3), np.array([world * box[0], box[1], box[2]]), rank, world,
                                               uid, dt=PHYS["dt"], rc=PHYS["rc"], skin=PHYS["skin"], sigma=PHYS["sigma"],
                                               epsilon=PHYS["epsilon"], cappingDistance=PHYS["cap"],
                                               maxNeighbors=PHYS["max_neigh"], langevin=True, zeta=PHYS["zeta"],
                                               temperature=PHYS["temperature"], seed=PHYS["seed"], cuts=cuts, **extra)
        return api.MolecularDynamics(a, sub, dt=PHYS["dt"], rc=PHYS["rc"], skin=PHYS["skin"], sigma=PHYS["sigma"],
                                     epsilon=PHYS["epsilon"], cappingDistance=PHYS["cap"],
                                     maxNeighbors=TETRAMER["max_neigh"] if tetramer else PHYS["max_neigh"], langevin=True,
                                     zeta=PHYS["zeta"], temperature=PHYS["temperature"], seed=PHYS["seed"],
                                     cellSort=not tetramer, fullList=int(args.full_list), **extra)

    md = make_md(atoms)
    md.run(args.equil, stream=stream)       # untimed: melt the lattice
    md.run(max(args.warmup, 3), stream=stream)  # warm-up

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            import torch.distributed as dist

            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    launches0 = api.launch_count()
    sampler.start()
    profile_range = bool(os.environ.get("MRMD_PROFILE_RANGE"))  # ncu --profile-from-start off: timed region only
    if profile_range:
        torch.cuda.profiler.start()
    ev0.record()
    stats = md.run(args.steps, timeForceKernel=True, stream=stream)
    ev1.record()
    if profile_range:
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    barrier()
    clocks = sampler.stop()
    launches = api.launch_count() - launches0
    ms = ev0.elapsed_time(ev1)
    t = torch.tensor([ms, float(stats["pairInteractions"]), float(launches), float(n)], dtype=torch.float64, device="cuda")
    total_atoms = float(n)
    if world > 1:
        import torch.distributed as dist

        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        # the slab driver already reports the pair interactions summed over the ranks
        ms_max, pairs_total = float(tmax[0]), (float(stats["pairInteractions"]) if slab_mode else float(tsum[1]))
        total_atoms = float(tsum[3])
    else:
        ms_max, pairs_total = ms, float(stats["pairInteractions"])
    value = total_atoms * args.steps / (ms_max * 1e-3)

    # roofline of the dominant kernel (LJ force): algorithmic bytes 60 N + 52 P_stored per launch
    peak, peak_src = measured_peak()
    # the contract figure is the half-list one whichever variant is timed (a full list stores every pair twice)
    stored_half = stats["storedPairs"] / (2.0 if args.full_list else 1.0)
    algo_bytes = 60.0 * n * args.steps + 52.0 * stored_half
    kernel_name = "ljForceTiledKernel" if args.full_list == 2 else "ljForceKernel"
    if adress:
        # SURVEY 8d, K14 with a = 1 atom per molecule: 84 M + 56 M a + 12 P_mol + (64 + 56 a) P_act
        algo_bytes = 140.0 * n * args.steps + 12.0 * stored_half + 120.0 * stats["activePairs"]
        kernel_name = "adressForceTiledKernel" if args.full_list == 2 else "adressForceKernel"
    if tetramer:
        # the same formula with a = 4 and M = n / 4 molecules: 84 M + 56 M a + 12 P_mol + (64 + 56 a) P_act
        algo_bytes = (84.0 + 224.0) * (n / 4) * args.steps + 12.0 * stats["storedPairs"] + 288.0 * stats["activePairs"]
        kernel_name = "adressActiveMoleculesKernel + adressForceLanes4Kernel"
    force_ms = stats["forceKernelMs"]
    n_kernel, kernel_rank, stored_kernel = n, rank, stats["storedPairs"]
    if world > 1:
        # report the rank whose force kernel ran longest (AdResS slabs carry very different work)
        import torch.distributed as dist

        mine = torch.tensor([force_ms, algo_bytes, float(stats["numLocal"]), float(stats["storedPairs"])],
                            dtype=torch.float64, device="cuda")
        every = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(every, mine)
        kernel_rank = int(np.argmax([float(e[0]) for e in every]))
        force_ms, algo_bytes = float(every[kernel_rank][0]), float(every[kernel_rank][1])
        n_kernel, stored_kernel = int(every[kernel_rank][2]), float(every[kernel_rank][3])
    achieved = algo_bytes / (force_ms * 1e-3) / 1e9 if force_ms > 0 else None
    traffic = ncu_traffic()
    roofline = {
        "bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": (achieved / peak) if achieved else None, "peak_source": peak_src,
        "traffic": traffic["dram_bytes_per_launch"] if (traffic and not adress and n == 1000000) else None,
        "algorithmic_bytes_per_launch": algo_bytes / args.steps,
        "kernel_ms_per_launch": force_ms / args.steps, "kernel_share_of_step": force_ms / ms_max,
        "stored_pairs_per_atom": stored_kernel / args.steps / max(n_kernel, 1), "rank": kernel_rank,
    }

    # end to end through the C ABI with HOST buffers: per step H2D pos+vel, one step, D2H pos+vel+scalars
    e2e = None
    if not args.no_e2e and slab_mode:
        # same contract on every rank's slab: the number of resident atoms changes when atoms migrate, so the
        # host mirror is re-sized from the step's own statistics
        cap = int(1.25 * n) + 1024
        hpos, hvel = api.PinnedBuffer((cap, 3)), api.PinnedBuffer((cap, 3))
        nl = md.run(0)["numLocal"]
        hpos.array[:nl] = atoms.get("pos")[:nl]
        hvel.array[:nl] = atoms.get("vel")[:nl]

        def host_step(nl):
            atoms.write_ptr("pos", hpos.ptr, 0, nl, 3, 1, api.HOST, stream)
            atoms.write_ptr("vel", hvel.ptr, 0, nl, 3, 1, api.HOST, stream)
            nl = md.run(1, stream=stream)["numLocal"]
            atoms.read_ptr("pos", hpos.ptr, 0, nl, 3, 1, api.HOST, stream)
            atoms.read_ptr("vel", hvel.ptr, 0, nl, 3, 1, api.HOST, stream)
            return nl

        for _ in range(3):
            nl = host_step(nl)
        k = max(1, min(args.e2e_steps, args.steps))
        barrier()
        t0 = time.perf_counter()
        for _ in range(k):
            nl = host_step(nl)
        barrier()
        dt = time.perf_counter() - t0
    elif not args.no_e2e:
        hpos, hvel, hsc = api.PinnedBuffer((n, 3)), api.PinnedBuffer((n, 3)), api.PinnedBuffer((3,))
        hpos.array[:] = atoms.get("pos")[:n]
        hvel.array[:] = atoms.get("vel")[:n]
        md.run_host(3, hpos.ptr, hvel.ptr, hsc.ptr, stream=stream)
        k = max(1, min(args.e2e_steps, args.steps))
        barrier()
        t0 = time.perf_counter()
        md.run_host(k, hpos.ptr, hvel.ptr, hsc.ptr, stream=stream)
        barrier()
        dt = time.perf_counter() - t0
    if not args.no_e2e:
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            import torch.distributed as dist

            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": total_atoms * k / float(tt[0]), "unit": "atom-steps/s",
               "h2d_bytes_per_step": int(48 * total_atoms / world), "d2h_bytes_per_step": int(48 * total_atoms / world) + 24,
               "steps": k,
               "path": ("per rank: pinned host pos+vel of the resident atoms -> device (mrmd_b200_atoms_write), one "
                        "mrmd_b200_slab_run step, pos+vel back (mrmd_b200_atoms_read)") if slab_mode else
                       "mrmd_b200_md_run_host: pinned host pos+vel -> device, one step, pos+vel+{E,virial,maxDisp} back"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = host_threads()
        cpos, cvel = atoms.get("pos")[:n], atoms.get("vel")[:n]
        t0 = time.perf_counter()
        probe, omd = (cpu_loop_tetramer if tetramer else (cpu_loop_adress if adress else cpu_loop))(cpos, cvel, box, 3, 1, threads)
        per_step = probe["seconds"] / 3
        more = int(max(3, min(200, (args.cpu_seconds - (time.perf_counter() - t0)) / max(per_step, 1e-6))))
        res = omd.run(more)
        cpu = {"value": n * res["steps"] / res["seconds"], "unit": "atom-steps/s", "cores": threads, "kind": "port",
               "sample": f"{res['steps']} steps of the same {n}-atom state (downloaded from the GPU after the timed "
                         f"region), {res['rebuilds']} rebuilds, OpenMP restatement of the reference "
                         f"{'AdResS ' if adress else ''}path",
               "pair_interactions_per_s": res["pairInteractions"] / res["seconds"]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "atom-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, n),
            "pair_interactions_per_s": pairs_total / (ms_max * 1e-3),
            "rebuild_interval_steps": args.steps / max(stats["rebuilds"], 1),
            "ghosts_per_gpu": stats["numGhost"], "energy_per_atom": stats["energy"] / (n * (world if slab_mode else 1)),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
        }
        if cuts is not None:
            line["config"]["slab_cuts"] = [float(c) for c in cuts]
            line["config"]["parallelism"] += "; cost-balanced slab widths (narrow over AT + HY, wide over CG)"
        emit(line)
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line):
    """the one JSON line goes to the process's original stdout"""
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    global _REAL_STDOUT
    args = parse_args()
    # libraries (NCCL_DEBUG=VERSION, torchrun banners) may print to fd 1: keep stdout for the JSON line only
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
