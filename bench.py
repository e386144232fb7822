#!/usr/bin/env python
"""bench.py -- atom-steps/s (and pair-interactions/s) of the MRMD force + neighbour hot path on B200.

  python bench.py --gpus N --steps K --warmup W            our arm (hand-written sm_100a kernels via the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  the reference's CPU path (OpenMP restatement: the
                                                           reference itself needs Kokkos/Cabana, unbuildable offline)

A "step" is one MD time step of the workload (pre-force integrate, ghost refresh or neighbour rebuild, force
zero, LJ force, ghost fold-back, post-force integrate).  One JSON line on stdout (rank 0).

The headline line is BASELINE.json configs[1] (Lennard-Jones NVT, 1M atoms per GPU).  The same invocation also runs, as
nested blocks of that line (each with its own value / ms_per_step / roofline / e2e / cpu_baseline):
  N = 1      "adress_config3"    configs[2]: LJ / ideal-gas AdResS slab, 8M atoms, thermodynamic force
             "lj_config1"        configs[0]: examples/02 size (4096 atoms), latency bound
  N = 2, 4   "tetramer_config4"  configs[3]: AdResS LJ tetramers, 16.4M atoms, x-slabs
  N = 8      "adress_config5"    configs[4]: 64M-atom LJ-AdResS slab over 8 cost-balanced x-slabs
  N > 1      "multi_gpu_parity"  x-slab run == single-GPU run of the same system (LJ, AdResS and tetramers, Langevin on)
--only-headline skips them; --workload / --side / --side-x / --balance select a single custom case instead.
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
ADRESS_THERMO = dict(targetDensity=0.512, binWidth=0.25, modulation=2.0, sampleInterval=10, updateInterval=1000,
                     sigma=2.0, range=2.0)
TETRAMER = dict(constraint_iterations=3, bond_length=1.0, max_neigh=40)
SPACING = {"lj": 1.25, "adress": 1.25, "tetramer": 1.98425}


def parse_args(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=2000)
    p.add_argument("--warmup", type=int, default=300)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--side", type=int, default=None, help="lattice sites per box edge per GPU (100 -> 1M atoms)")
    p.add_argument("--side-x", type=int, default=None, help="lattice sites per GPU along x if different from --side "
                   "(configs[4]: --workload adress --side 400 --side-x 50 --gpus 8 = 64M atoms in a 500^3 box)")
    p.add_argument("--equil", type=int, default=300, help="untimed equilibration steps that melt the lattice")
    p.add_argument("--full-list", type=int, default=2, help="0 half list, 1 full list, 2 tiled periodic full list (fast path)")
    p.add_argument("--e2e-steps", type=int, default=100)
    p.add_argument("--cpu-seconds", type=float, default=20.0, help="budget of the cpu_baseline leg")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--only-headline", action="store_true", help="skip the nested config blocks and the parity check")
    p.add_argument("--replicas", action="store_true", help="N>1: independent periodic replicas instead of x-slabs")
    p.add_argument("--balance", action="store_true", help="adress workload on N>1 GPUs: cost-balanced slab widths")
    p.add_argument("--force-cost-ratio", type=float, default=4.5,
                   help="--balance: force-kernel time / rest of the step, per atom of the AT + HY region")
    p.add_argument("--workload", default=None, choices=["lj", "adress", "tetramer"],
                   help="a single custom case instead of the default set. lj: configs[1]; adress: configs[2]/[4] physics "
                        "(LJ / ideal-gas AdResS slab with thermodynamic force, one molecule per atom; --side 200 = 8M atoms "
                        "per GPU); tetramer: configs[3] physics (AdResS LJ tetramers, spherical region, SHAKE / RATTLE; "
                        "--side = molecules per edge, 160 -> 16.4M atoms)")
    args = p.parse_args(argv)
    args.custom = args.workload is not None or args.side is not None or args.side_x is not None
    if args.workload is None:
        args.workload = "lj"
    if args.side is None:
        args.side = 100
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

    def sample_once(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
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

    def run(self):
        if self.nv is None:
            return
        while not self._stop_evt.is_set():
            self.sample_once()
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


def ncu_profile(kernel, n_atoms):
    """per-launch ncu figures of a dominant kernel at this size from the committed capture (profiles/kernel_ncu.json,
    written by profiles/ncu_summary.py from an `ncu --set full` report of this bench), or None"""
    path = os.path.join(ROOT, "profiles", "kernel_ncu.json")
    if os.path.exists(path):
        try:
            table = json.load(open(path))
            exact = table.get(f"{kernel}@{int(n_atoms)}")
            if exact is not None:
                return exact
            # no capture at this size (e.g. a cost-balanced slab): the capture of the same kernel at the nearest size, its
            # byte counts scaled by the number of atoms (the pipe / issue fractions are intensive)
            sizes = [(int(k.split("@")[1]), v) for k, v in table.items() if k.split("@")[0] == kernel]
            if sizes and n_atoms > 0:
                ref_atoms, ref = min(sizes, key=lambda kv: abs(kv[0] - n_atoms))
                if not 0.25 <= float(n_atoms) / ref_atoms <= 4.0:
                    return None  # another regime (e.g. the 4096-atom case is latency bound)
                scaled = dict(ref)
                for key in ("dram_bytes_read", "dram_bytes_write", "dram_bytes_per_launch", "warp_instructions"):
                    if key in scaled:
                        scaled[key] = scaled[key] * float(n_atoms) / ref_atoms
                scaled.pop("duration_us_under_ncu", None)
                scaled["source"] = f"{ref.get('source')}; byte counts scaled from {ref_atoms} to {int(n_atoms)} atoms"
                return scaled
        except Exception:
            pass
    return None


# ---- the CPU leg (oracle/: checker and baseline only) -----------------------------------------------------------
def cpu_md(workload, pos, vel, box, threads):
    """the reference path's OpenMP restatement on the host cores for one workload"""
    from oracle import pyoracle as orc
    from oracle.md_loop import OracleAdressMD, OracleMD

    orc.build()
    orc.lib().or_set_threads(threads)
    common = dict(dt=PHYS["dt"], rc=PHYS["rc"], skin=PHYS["skin"], sigma=PHYS["sigma"], epsilon=PHYS["epsilon"],
                  cap=PHYS["cap"], langevin=True, zeta=PHYS["zeta"], temperature=PHYS["temperature"], seed=PHYS["seed"])
    if workload == "lj":
        return OracleMD(pos, vel, box, max_neigh=PHYS["max_neigh"], cell_sort=True, **common)
    if workload == "adress":
        lx = float(box[0])
        weight = orc.make_weight(orc.WEIGHT_SLAB, [lx / 2, box[1] / 2, box[2] / 2], 0.2 * lx, 0.1 * lx, 1)
        return OracleAdressMD(pos, vel, box, weight, max_neigh=PHYS["max_neigh"], thermo=ADRESS_THERMO, **common)
    radius, hybrid = tetramer_region(box)
    weight = orc.make_weight(orc.WEIGHT_SPHERICAL, 0.5 * np.asarray(box), radius, hybrid, 2)
    return OracleAdressMD(pos, vel, box, weight, max_neigh=TETRAMER["max_neigh"], atoms_per_mol=4,
                          constraint_iterations=TETRAMER["constraint_iterations"], bond_length=TETRAMER["bond_length"],
                          **common)


def tetramer_region(box):
    """config 4: spherical region, centre = box centre, R = 60 and h = 30 in the 317.48 box (scaled with the box)"""
    return 60.0 / 317.48 * float(box[1]), 30.0 / 317.48 * float(box[1])


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def make_system(workload, side, side_x=None, seed=None):
    from mrmd_b200.workloads import lattice_system, tetramer_system

    seed = PHYS["seed"] if seed is None else seed
    if workload == "tetramer":
        return tetramer_system(side, seed=seed, n_side_x=side_x)
    return lattice_system(side, seed=seed, n_side_x=side_x)


def reference_case(workload, side, steps, warmup, equil, threads, budget_s):
    """One workload on the host cores: the same lattice system melted by the same number of equilibration steps when
    that fits the time budget, otherwise a smaller periodic system at the same state point (atom-steps/s is size
    normalised).  Returns the block of the JSON line."""
    probe_side = 20 if workload == "tetramer" else 32
    pos, vel, box = make_system(workload, probe_side)
    md = cpu_md(workload, pos, vel, box, threads)
    md.run(2)
    probe = md.run(6)
    rate = len(pos) * probe["steps"] / probe["seconds"]
    total_steps = max(equil + warmup + steps, 1)
    budget_atoms = rate * budget_s / total_steps / (4.0 if workload == "tetramer" else 1.0)
    run_side = int(max(16, min(side, np.floor(budget_atoms ** (1.0 / 3.0)))))
    pos, vel, box = make_system(workload, run_side)
    n = len(pos)
    md = cpu_md(workload, pos, vel, box, threads)
    md.run(equil)
    md.run(warmup)
    res = md.run(steps)
    value = n * res["steps"] / res["seconds"]
    what = (f"the {side}^3 workload itself" if run_side == side else
            f"a smaller periodic system ({run_side}^3 sites, same rho/T/dt/skin as the {side}^3 workload)")
    sample = (f"{what}: {n} atoms, {equil} equilibration + {warmup} warm-up + {steps} timed steps, "
              f"{res['rebuilds']} neighbour rebuilds in the timed steps")
    return {
        "value": value, "unit": "atom-steps/s", "ms_per_step": 1e3 * res["seconds"] / max(res["steps"], 1),
        "pair_interactions_per_s": res["pairInteractions"] / res["seconds"],
        "cpu_baseline": {"value": value, "unit": "atom-steps/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "atom-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "atoms": n,
    }


def run_reference(args):
    """--impl reference: rank 0 only; every step a bounded sample so that the run ends within a few minutes."""
    rank, _, world = env_rank()
    if rank != 0:
        return
    threads = host_threads()
    head = reference_case(args.workload, args.side, args.steps, args.warmup, args.equil, threads, 120.0)
    atoms_per_mol = 4 if args.workload == "tetramer" else 1
    line = {
        "impl": "reference", "metric": METRIC, "value": head["value"], "unit": "atom-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": head["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.workload, args.side, args.side_x, args.equil, args.full_list, args.gpus, False,
                                  args.side ** 3 * atoms_per_mol),
        "pair_interactions_per_s": head["pair_interactions_per_s"],
        "cpu_baseline": head["cpu_baseline"], "e2e": head["e2e"],
        "note": "OpenMP restatement of the reference path (Kokkos 4.7.1 / Cabana 0.7 un-vendored: the reference "
                "cannot be built offline)",
    }
    if not args.custom and not args.only_headline:
        # the nested workloads of the B200 arm on the host cores, bounded samples of ~40 s each
        nested = {1: [("adress_config3", "adress", 200)], 2: [("tetramer_config4", "tetramer", 160)],
                  4: [("tetramer_config4", "tetramer", 160)], 8: [("adress_config5", "adress", 400)]}.get(args.gpus, [])
        for name, wl, side in nested:
            try:
                line[name] = reference_case(wl, side, args.steps, args.warmup, min(args.equil, 100), threads, 40.0)
            except Exception as exc:  # the headline must survive a failing side case
                line[name] = {"error": f"{type(exc).__name__}: {exc}"}
    emit(line)


def workload_config(workload, side, side_x, equil, full_list, gpus, replicas, n_atoms_per_gpu, balanced=False):
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
           "on the six bonds (SHAKE / RATTLE), Langevin gamma=20 T=1.5, dt=0.002, skin 0.1, tiled Verlet list of the "
           "molecules' centres of mass, maxNeighbors 40")
    sp = SPACING[workload]
    sx = side_x or side
    if gpus == 1:
        par = "1 GPU"
    elif replicas:
        par = f"{gpus} independent periodic replicas, one per GPU"
    else:
        par = (f"{gpus} x-slabs of one {gpus * sx * sp:g} x {side * sp:g} x {side * sp:g} box, one process per GPU; per step: "
               "face-atom positions stored into the neighbours' CUDA-IPC mapped buffers over NVLink (peer stores + sequence "
               "flags, no NCCL call), rebuild decision = all-gather of max |dx|^2 through the same peer buffers; at a "
               "rebuild: atoms migrate and halo lists are re-selected; no reverse force halo (full list); NCCL bootstraps "
               "the mapping and carries the read-out reductions")
        if balanced:
            par += "; cost-balanced slab widths (narrow over AT + HY, wide over CG)"
    return {
        "workload": {"lj": lj, "adress": ad, "tetramer": tet}[workload],
        "atoms_per_gpu": n_atoms_per_gpu, "box_per_gpu": [sx * sp, side * sp, side * sp], "equilibration_steps": equil,
        "list": {0: "half (reference semantics, fp64 RED scatter)", 1: "full (generic gather kernel)",
                 2: "full, periodic tiles staged in shared memory (mrmd_b200_verlet_build_periodic)"}[full_list],
        "l2": "inputs larger than L2 (per step: 104 B/atom state + 2 B x 64-slot neighbour rows > 126 MB at 1M atoms); "
              "no explicit flush",
        "parallelism": par,
    }


class Ctx:
    """process-wide handles of the B200 arm"""

    def __init__(self):
        import torch

        self.torch = torch
        self.rank, self.local_rank, self.world = env_rank()
        torch.cuda.set_device(self.local_rank)
        self.dist = None
        self.numa_bound = False
        if self.world > 1:
            import torch.distributed as dist
            from mrmd_b200.slabs import bind_host_to_gpu

            # one process per GPU: host threads and pinned buffers on the GPU's NUMA node (the cpu_baseline leg runs at
            # N = 1 only and keeps all cores)
            self.numa_bound = bind_host_to_gpu(self.local_rank) is not None

            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.dist = dist
        from mrmd_b200 import api

        self.api = api
        api.L().mrmd_b200_set_device(self.local_rank)
        self.stream = torch.cuda.current_stream().cuda_stream

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, values, op):
        t = self.torch.tensor(values, dtype=self.torch.float64, device="cuda")
        if self.dist is not None:
            self.dist.all_reduce(t, op={"max": self.dist.ReduceOp.MAX, "sum": self.dist.ReduceOp.SUM}[op])
        return [float(x) for x in t]


def run_case(ctx, workload, side, side_x, steps, warmup, equil, full_list=2, balance=False, force_cost_ratio=4.5,
             replicas=False, e2e_steps=100, cpu_seconds=20.0, want_cpu=True, want_e2e=True, clocks=True):
    """One workload on the B200 arm; returns the block (rank 0) -- collective over all ranks."""
    torch, api, rank, world = ctx.torch, ctx.api, ctx.rank, ctx.world
    stream = ctx.stream
    slab_mode = world > 1 and not replicas
    tetramer = workload == "tetramer"
    adress = workload in ("adress", "tetramer")
    apm = 4 if tetramer else 1
    sites_x = side_x or side
    spacing = SPACING[workload]
    global_lx = world * sites_x * spacing if slab_mode else sites_x * spacing
    cuts = None
    my_sites_x, x_offset = sites_x, (rank * sites_x * spacing if slab_mode else 0.0)
    if slab_mode and workload == "adress" and balance:
        # cost-balanced slabs: the AT + HY region [0.3 Lx, 0.7 Lx] carries the force kernel, the
        # coarse-grained rest only streams; boundaries on lattice planes so that every rank builds its own atoms
        from mrmd_b200 import slabs

        cuts = slabs.balanced_cuts([0.0], [global_lx], world, 0.3 * global_lx, 0.7 * global_lx, force_cost_ratio,
                                   quantum=spacing, min_width=2 * (PHYS["rc"] + PHYS["skin"]))
        my_sites_x = int(round((cuts[rank + 1] - cuts[rank]) / spacing))
        x_offset = float(cuts[rank])
    pos, vel, box = make_system(workload, side, side_x=my_sites_x, seed=PHYS["seed"] + rank)
    box = np.array([sites_x * spacing, box[1], box[2]])  # the equal-width slab: world * box[0] is the global length
    n = len(pos)
    sub = api.Subdomain([0, 0, 0], box, PHYS["rc"] + PHYS["skin"])
    ids = None
    if slab_mode:
        # weak scaling over x-slabs: one global box of world * sites_x x side x side sites, rank r owns slab r
        pos = pos + np.array([x_offset, 0.0, 0.0])
        first = int(round(x_offset / spacing)) * side * side * apm  # global ids: lattice order of the global system
        ids = first + np.arange(n, dtype=np.int64)
    atoms = api.Atoms.from_arrays(pos, vel, mass=1.0, relativeMass=1.0 / apm, ids=ids)

    extra = {}
    global_box = np.array([global_lx, box[1], box[2]])
    if tetramer:
        radius, hybrid = tetramer_region(global_box)
        extra = dict(adress=True, weight=api.Spherical(0.5 * global_box, radius, hybrid, 2), doShift=True, atomsPerMolecule=4,
                     numConstraintIterations=TETRAMER["constraint_iterations"], bondLength=TETRAMER["bond_length"])
    elif adress:
        # configs[2] / configs[4] physics: Slab(centre of the global box, AT diameter 0.2 Lx, HY width 0.1 Lx, nu = 1),
        # LJ_IdealGas(cap 0.7, rc 2.5, shift on), ThermodynamicForce(rho 0.512, bin 0.25, modulation 2; sample every
        # 10 steps, update(sigma 2, range 2) every 1000 steps), one molecule per atom
        extra = dict(adress=True, weight=api.Slab(0.5 * global_box, 0.2 * global_lx, 0.1 * global_lx, 1), doShift=True,
                     thermo=dict(ADRESS_THERMO))
    common = dict(dt=PHYS["dt"], rc=PHYS["rc"], skin=PHYS["skin"], sigma=PHYS["sigma"], epsilon=PHYS["epsilon"],
                  cappingDistance=PHYS["cap"], maxNeighbors=TETRAMER["max_neigh"] if tetramer else PHYS["max_neigh"],
                  langevin=True, zeta=PHYS["zeta"], temperature=PHYS["temperature"], seed=PHYS["seed"])
    if slab_mode:
        from mrmd_b200 import slabs

        uid = slabs.broadcast_unique_id(rank)
        md = slabs.SlabMolecularDynamics(atoms, np.zeros(3), global_box, rank, world, uid, cuts=cuts, **common, **extra)
    else:
        md = api.MolecularDynamics(atoms, sub, cellSort=not (tetramer and full_list == 0), fullList=int(full_list), **common,
                                   **extra)

    md.run(equil, stream=stream)            # untimed: melt the lattice
    md.run(max(warmup, 3), stream=stream)   # warm-up

    def timed(nsteps, time_kernel):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx.barrier()
        ev0.record()
        st = md.run(nsteps, timeForceKernel=time_kernel, stream=stream)
        ev1.record()
        ctx.barrier()
        return ev0.elapsed_time(ev1), st

    sampler = ClockSampler(ctx.local_rank) if clocks else None
    launches0 = api.launch_count()
    if sampler is not None:
        sampler.start()
    profile_range = bool(os.environ.get("MRMD_PROFILE_RANGE"))  # ncu --profile-from-start off: timed region only
    if profile_range:
        torch.cuda.profiler.start()
    ms, stats = timed(steps, True)
    if profile_range:
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    if sampler is not None and sampler.nv is not None and not sampler.samples:
        sampler.sample_once()  # a timed region shorter than the sampling period
    clock_info = sampler.stop() if sampler is not None else None
    launches = api.launch_count() - launches0
    ms_max, = ctx.reduce([ms], "max")
    psum, nsum = ctx.reduce([float(stats["pairInteractions"]), float(n)], "sum")
    pairs_total = float(stats["pairInteractions"]) if slab_mode else psum  # the slab driver reports the global sum
    total_atoms = nsum
    value = total_atoms * steps / (ms_max * 1e-3)

    # the same steps with energy and virial reduced on EVERY step, as LennardJones::apply does (the headline reduces them
    # on the last step of a run only: they are not observable earlier)
    md.setEnergyEveryStep(True)
    md.run(2, stream=stream)
    ms_e, _ = timed(steps, False)
    md.setEnergyEveryStep(False)
    ms_e_max, = ctx.reduce([ms_e], "max")

    # roofline of the dominant kernel: algorithmic bytes per launch of the SURVEY 8d contract (half-list figure
    # whichever variant is timed: a full list stores every pair twice)
    peak, peak_src = measured_peak()
    stored_half = stats["storedPairs"] / (2.0 if full_list else 1.0)
    algo_bytes = 60.0 * n * steps + 52.0 * stored_half
    kernel_name = "ljForceTiledKernel" if full_list == 2 else "ljForceKernel"
    if adress:
        # K14 with a = 1 atom per molecule: 84 M + 56 M a + 12 P_mol + (64 + 56 a) P_act
        algo_bytes = 140.0 * n * steps + 12.0 * stored_half + 120.0 * stats["activePairs"]
        kernel_name = "adressForceTiledKernel" if full_list == 2 else "adressForceKernel"
    if tetramer:
        # the same formula with a = 4 and M = n / 4 molecules
        algo_bytes = (84.0 + 224.0) * (n / 4) * steps + 12.0 * stored_half + 288.0 * stats["activePairs"]
        kernel_name = ("moleculeForceTiledKernel" if full_list == 2 else
                       "adressActiveMoleculesKernel + adressForceLanes4Kernel")
    force_ms = stats["forceKernelMs"]
    n_kernel, kernel_rank, stored_kernel = n, rank, stats["storedPairs"]
    if world > 1:
        # report the rank whose force kernel ran longest (AdResS slabs carry very different work)
        mine = torch.tensor([force_ms, algo_bytes, float(stats["numLocal"]), float(stats["storedPairs"])],
                            dtype=torch.float64, device="cuda")
        every = [torch.zeros_like(mine) for _ in range(world)]
        ctx.dist.all_gather(every, mine)
        kernel_rank = int(np.argmax([float(e[0]) for e in every]))
        force_ms, algo_bytes = float(every[kernel_rank][0]), float(every[kernel_rank][1])
        n_kernel, stored_kernel = int(every[kernel_rank][2]), float(every[kernel_rank][3])
    achieved = algo_bytes / (force_ms * 1e-3) / 1e9 if force_ms > 0 else None
    prof = ncu_profile(kernel_name, n_kernel if world == 1 else n) or {}
    traffic = prof.get("dram_bytes_per_launch")
    per_launch_s = force_ms * 1e-3 / steps if force_ms > 0 else None
    roofline = {
        # contract figure: SURVEY 8(d) algorithmic bytes / measured kernel time against the measured HBM copy peak.  The
        # kernel does not move those bytes (positions are staged once per tile, 2-byte list entries), so the figure can
        # exceed 1: the DRAM traffic ncu measures is `traffic`, and the units that limit the kernel are the shared-memory
        # (L1TEX) data path of the slot gathers, the FP64 pipe and warp issue (l1tex_frac / pipe_frac / issue_frac)
        "bound": "fp64", "limiter": "on-chip: shared-memory gathers of the staged partners (L1TEX), FP64 pipe, warp issue; "
                                    "frac is the SURVEY 8(d) contract figure (reference-layout bytes / kernel time / HBM "
                                    "peak) and exceeds the DRAM share the kernel really uses (dram_frac)",
        "kernel": kernel_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": (achieved / peak) if achieved else None, "contract_hbm_frac": (achieved / peak) if achieved else None,
        "peak_source": peak_src, "traffic": traffic,
        "dram_frac": (traffic / per_launch_s / 1e9 / peak) if (traffic and per_launch_s) else None,
        "pipe_frac": prof.get("fp64_pipe_frac"), "issue_frac": prof.get("issue_active_frac"),
        "l1tex_frac": prof.get("l1tex_throughput_frac"),
        "ncu_source": prof.get("source"),
        "algorithmic_bytes_per_launch": algo_bytes / steps,
        "kernel_ms_per_launch": force_ms / steps, "kernel_share_of_step": force_ms / ms_max,
        "stored_pairs_per_atom": stored_kernel / steps / max(n_kernel, 1), "rank": kernel_rank,
    }

    # end to end through the C ABI with HOST buffers: per step H2D pos+vel, one step, D2H pos+vel+scalars
    e2e = None
    if want_e2e:
        cap = int(1.25 * n) + 1024 if slab_mode else n
        hpos, hvel, hsc = api.PinnedBuffer((cap, 3)), api.PinnedBuffer((cap, 3)), api.PinnedBuffer((4,))
        nl = md.run(0)["numLocal"]
        hpos.array[:nl] = atoms.get("pos")[:nl]
        hvel.array[:nl] = atoms.get("vel")[:nl]
        k = max(1, min(e2e_steps, steps))

        def host_run(steps_per_call, calls):
            t = 0.0
            ctx.barrier()
            t0 = time.perf_counter()
            for _ in range(calls):
                md.run_host(steps_per_call, hpos.ptr, hvel.ptr, hsc.ptr, stream=stream)
            ctx.barrier()
            t = time.perf_counter() - t0
            return ctx.reduce([t], "max")[0]

        host_run(3, 1)
        t_pipe = host_run(k, 1)       # one call, k steps: the upload of step i+1 trails the download of step i
        t_single = host_run(1, k)     # k calls of one step each: no overlap across steps (the host may touch the buffers)
        per_gpu_bytes = int(48 * total_atoms / world)
        e2e = {"value": total_atoms * k / t_pipe, "unit": "atom-steps/s",
               "h2d_bytes_per_step": per_gpu_bytes, "d2h_bytes_per_step": per_gpu_bytes + 32, "steps": k,
               "value_one_step_per_call": total_atoms * k / t_single,
               "path": ("mrmd_b200_slab_run_host on every rank" if slab_mode else "mrmd_b200_md_run_host") +
                       ": pinned host pos+vel of the resident atoms -> device, one step, pos+vel+{E,virial,maxDisp} "
                       "back; chunked copies on two copy streams, step i+1's upload trails step i's download; "
                       "bytes are per GPU; value_one_step_per_call = the same with one call per step (no overlap "
                       "across steps)",
               "host_numa_bound": bool(ctx.numa_bound)}
        del hpos, hvel, hsc

    cpu = None
    if rank == 0 and world == 1 and want_cpu:
        threads = host_threads()
        cpos, cvel = atoms.get("pos")[:n], atoms.get("vel")[:n]
        t0 = time.perf_counter()
        omd = cpu_md(workload, cpos, cvel, box, threads)
        probe = omd.run(3)
        per_step = probe["seconds"] / 3
        more = int(max(2, min(200, (cpu_seconds - (time.perf_counter() - t0)) / max(per_step, 1e-6))))
        res = omd.run(more)
        cpu = {"value": n * res["steps"] / res["seconds"], "unit": "atom-steps/s", "cores": threads, "kind": "port",
               "sample": f"{res['steps']} steps of the same {n}-atom state (downloaded from the GPU after the timed "
                         f"region), {res['rebuilds']} rebuilds, OpenMP restatement of the reference "
                         f"{'AdResS ' if adress else ''}path",
               "pair_interactions_per_s": res["pairInteractions"] / res["seconds"]}
        del omd

    block = {
        "value": value, "unit": "atom-steps/s", "ms_per_step": ms_max / steps, "steps": steps, "warmup": max(warmup, 3),
        "atoms_total": int(total_atoms),
        "energy_reduction": "last step of the run only (md_config.energyEveryStep = 0); force and pair count every step",
        "value_energy_every_step": total_atoms * steps / (ms_e_max * 1e-3),
        "ms_per_step_energy_every_step": ms_e_max / steps,
        "config": workload_config(workload, side, side_x, equil, full_list, world, replicas, n, balanced=cuts is not None),
        "pair_interactions_per_s": pairs_total / (ms_max * 1e-3),
        "rebuild_interval_steps": steps / max(stats["rebuilds"], 1),
        "ghosts_per_gpu": stats["numGhost"], "energy_per_atom": stats["energy"] / (n * (world if slab_mode else 1)),
        "clocks": clock_info, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
    }
    if cuts is not None:
        block["config"]["slab_cuts"] = [float(c) for c in cuts]
    md.close() if hasattr(md, "close") else None
    del md, atoms
    torch.cuda.empty_cache()
    return block


def run_b200(args):
    ctx = Ctx()
    world, rank = ctx.world, ctx.rank
    line = {"metric": METRIC}
    parity = None
    if world > 1 and not args.custom and not args.only_headline and not args.replicas:
        # before anything is timed: the x-slab data plane must reproduce the single-GPU run (Langevin on)
        from mrmd_b200 import slabs

        parity = {}
        for mode in ("lj", "adress", "tetramer"):
            try:
                parity[mode] = slabs.parity_check(rank, world, steps=40, mode=mode, langevin=True, stream=ctx.stream)
            except Exception as exc:
                parity[mode] = {"ok": False, "error": f"{type(exc).__name__}: {exc}"}
        parity["ok"] = all(v.get("ok") for v in parity.values())
    head = run_case(ctx, args.workload, args.side, args.side_x, args.steps, args.warmup, args.equil,
                    full_list=args.full_list, balance=args.balance, force_cost_ratio=args.force_cost_ratio,
                    replicas=args.replicas, e2e_steps=args.e2e_steps, cpu_seconds=args.cpu_seconds,
                    want_cpu=not args.no_cpu_baseline, want_e2e=not args.no_e2e)
    line.update({k: head[k] for k in ("value", "unit")})
    line.update({"n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": head["ms_per_step"],
                 "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic"})
    line.update({k: v for k, v in head.items() if k not in line})
    if parity is not None:
        line["multi_gpu_parity"] = parity
    if not args.custom and not args.only_headline and not args.replicas:
        nested = {
            1: [("adress_config3", dict(workload="adress", side=200, side_x=None, balance=False, cpu_seconds=12.0)),
                ("lj_config1", dict(workload="lj", side=16, side_x=None, balance=False, cpu_seconds=4.0,
                                    steps=max(args.steps, 2000), warmup=max(args.warmup, 200)))],
            2: [("tetramer_config4", dict(workload="tetramer", side=160, side_x=80, balance=False))],
            4: [("tetramer_config4", dict(workload="tetramer", side=160, side_x=40, balance=False))],
            8: [("adress_config5", dict(workload="adress", side=400, side_x=50, balance=True))],
        }.get(world, [])
        for name, kw in nested:
            try:
                line[name] = run_case(ctx, kw["workload"], kw["side"], kw["side_x"], kw.get("steps", args.steps),
                                      kw.get("warmup", args.warmup), args.equil, balance=kw["balance"],
                                      force_cost_ratio=args.force_cost_ratio, e2e_steps=min(args.e2e_steps, 20),
                                      cpu_seconds=kw.get("cpu_seconds", 10.0), want_cpu=not args.no_cpu_baseline,
                                      want_e2e=not args.no_e2e)
            except Exception as exc:  # the headline must survive a failing side case
                line[name] = {"error": f"{type(exc).__name__}: {exc}"}
    if rank == 0:
        emit(line)
    if ctx.dist is not None:
        ctx.dist.destroy_process_group()


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
