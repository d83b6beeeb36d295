#!/usr/bin/env python3
"""bench.py -- zone-updates/s of one full hydro + ion-RT step (BASELINE.json metric) on N B200s.

A "step" is one pass of the reference's main-loop body (main.c:519-669) over the whole grid:
[ion_radtransfer_3d sub-cycle; bvals_mhd] -> integrate_3d_ctu -> Userwork_in_loop -> new_dt -> bvals_mhd,
driven through the C ABI of libathena_b200.so.  Default workload: ioniz_sphere unigrid 256^3 (BASELINE.json
configs[2], the north-star target config: hydro + ion RT + static potential); `--workload blast` is configs[1]
(hydro-only isolation), `--workload ifront` is configs[0].  With --gpus N the grid is x3-slab decomposed, one
256x256x256 slab per GPU (weak scaling), ghost slabs over NCCL, timestep/sub-cycle scalars over all_reduce.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload ioniz_sphere|blast|ifront] [--n 256]
    python bench.py --impl reference ...      # the reference's own CPU build (oracle/_ref) on the host cores

Prints ONE JSON line (rank 0).  `value` = zones of all ranks x K / max-over-ranks device time with the state
resident in HBM; `e2e` = the same through host buffers (ConsS AoS upload + step + download per step);
`roofline` = the dominant kernel of the timed region, CUDA-event timed on its stream, against MEASURED_PEAKS.json.
"""
import argparse
import json
import os
import re
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

load_package()
from athena_b200.athinput import ParFile  # noqa: E402

DECKS = os.path.join(ROOT, "tests", "decks")
OUT = sys.stdout
WORKLOADS = {
    # name: (deck, problem/reference build, BASELINE.json config it corresponds to)
    "ioniz_sphere": ("ioniz_sphere.athinput", "ioniz_sphere", "configs[2] ioniz_sphere unigrid"),
    "blast": ("blast.athinput", "blast", "configs[1] 3D hydro blast, hydro-only"),
    "ifront": ("ifront.athinput", "ifront", "configs[0] ifront"),
}

# Compulsory HBM bytes per zone of each kernel given the pass structure (DESIGN.md "Kernels"), as
# (doubles with nvar multiplier, plain doubles, doubles only present with a static potential); x 8 B.
KERNEL_BYTES = {
    "k_first_x1": (4, 0, 2), "k_first_march<1>": (4, 0, 2), "k_first_march<2>": (4, 0, 2),
    "k_dhalf": (0, 5, 0),
    "k_correct<0>": (6, 1, 4), "k_correct<1>": (6, 1, 4), "k_correct<2>": (6, 1, 4),
    "k_final<0>": (3, 3, 0), "k_final<1>": (3, 3, 0), "k_final<2>": (3, 3, 0),
    "k_correct3": (15, 3, 5),    # Ul,Ur of three directions read and written, F of three directions read once, 3 eta written; U0 + 4 potential tables
    "k_update": (5, 0, 7),       # U read and written, final fluxes of three directions; 4 potential tables + 3 first-pass mass fluxes (d^{n+1/2})
    "k_ion_prep": (0, 11.5, 0), "k_ion_rates": (0, 11, 0), "k_ion_rates[first]": (0, 14, 0),   # [first]: + e_init, e_th_init, x_init written, sign states not read "k_ion_update": (0, 14, 0), "k_ion_rates<fused update>": (0, 19, 0),
    "k_max_speeds": (0, 5, 0),
    "k_cfinal": (12, 0, 6),      # first-pass Ul,Ur,F of three directions read once, final fluxes written once
}


def kernel_bytes_per_zone(name, nvar, grav):
    if name not in KERNEL_BYTES:
        return None
    a, b, c = KERNEL_BYTES[name]
    return 8.0 * (a * nvar + b + (c if grav else 0))


def step_algorithmic_bytes(workload, n_iter):
    """B_alg = B_hydro + n_iter*B_ion per zone-update (compulsory traffic, FP64, SoA)."""
    b_hydro = {"blast": 80.0, "ifront": 96.0, "ioniz_sphere": 128.0}[workload]
    return b_hydro + (n_iter * 112.0 if workload != "blast" else 0.0)


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def make_par(workload, n, nx3):
    deck, _, _ = WORKLOADS[workload]
    par = ParFile.par_open(os.path.join(DECKS, deck))
    par.par_cmdline(["job/num_domains=1", f"domain1/Nx1={n}", f"domain1/Nx2={n}", f"domain1/Nx3={nx3}"])
    if nx3 != n:    # weak scaling: the box grows along x3 with the GPU count, zones stay cubic
        x3min, x3max = par.par_getd("domain1", "x3min"), par.par_getd("domain1", "x3max")
        c, h = 0.5 * (x3min + x3max), 0.5 * (x3max - x3min) * nx3 / n
        par.par_cmdline([f"domain1/x3min={c - h!r}", f"domain1/x3max={c + h!r}"])
    return par


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.lines = []
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(index)],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.lines.append((time.time(), line.strip()))

    def summary(self, t0, t1):
        if self.p is not None:
            self.p.terminate()
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.lines:
            if ts < t0 - 0.15 or ts > t1 + 0.3:          # 100 ms sampling period: the sample that straddles the start counts
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1]))
            except Exception:
                continue
            for nme, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the reference's own binary (oracle/_ref/<cfg>/athena), serial (no MPI here)
# ---------------------------------------------------------------------------------------------------------
def reference_exe(workload):
    return os.path.join(ROOT, "oracle", "_ref", WORKLOADS[workload][1], "athena")


def run_reference_cpu(workload, n, nsteps, instances):
    """`instances` concurrent serial runs of the reference binary on an n^3 grid for nsteps steps.
    Returns (sum of zone-cycles/wall-second over instances, wall seconds, kind)."""
    exe = reference_exe(workload)
    deck = os.path.join(DECKS, WORKLOADS[workload][0])
    if not os.path.exists(exe):
        return run_port_cpu(workload, min(n, 32), max(2, min(nsteps, 3)))
    procs, tmps = [], []
    t0 = time.time()
    for _ in range(instances):
        tmp = tempfile.mkdtemp(prefix="ath_ref_")
        tmps.append(tmp)
        cmd = [exe, "-i", deck, "job/maxout=0", "job/num_domains=1", f"domain1/Nx1={n}", f"domain1/Nx2={n}", f"domain1/Nx3={n}", f"time/nlim={nsteps}"]
        procs.append(subprocess.Popen(cmd, cwd=tmp, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True))
    total = 0.0
    for p in procs:
        out, _ = p.communicate()
        m = re.search(r"zone-cycles/wall-second = (\S+)", out)
        if m:
            total += float(m.group(1))
    wall = time.time() - t0
    for tmp in tmps:
        subprocess.run(["rm", "-rf", tmp])
    return total, wall, "reference"


def run_dropin(workload, n, nsteps):
    """The UNMODIFIED reference driver linked against libathena_b200 (build/dropin/<problem>/athena) on
    the bench grid: the throughput a user of the reference sees after relinking -- the reference's own
    `zone-cycles/wall-second` line (main.c:722: main-loop wall clock, host `U` uploaded before the first step and
    downloaded for the final dump, coherence in between only where a writer fires).  None if the executable is absent."""
    prob = WORKLOADS[workload][1]
    exe = os.path.join(ROOT, "build", "dropin", prob, "athena")
    if not os.path.exists(exe):
        return None
    tmp = tempfile.mkdtemp(prefix="ath_dropin_")
    cmd = [exe, "-i", os.path.join(DECKS, WORKLOADS[workload][0]), "job/maxout=0", "job/num_domains=1",
           f"domain1/Nx1={n}", f"domain1/Nx2={n}", f"domain1/Nx3={n}", f"time/nlim={nsteps}"]
    t0 = time.time()
    env = dict(os.environ)
    env.setdefault("AB2_GRAV_THREADS", str(max(1, min(32, os.cpu_count() or 1))))    # prob/ioniz_sphere.c PlanetPot only reads file statics: re-entrant
    r = subprocess.run(cmd, cwd=tmp, capture_output=True, text=True, env=env)
    wall = time.time() - t0
    subprocess.run(["rm", "-rf", tmp])
    m = re.search(r"zone-cycles/wall-second = (\S+)", r.stdout)
    if r.returncode != 0 or not m:
        return {"error": (r.stdout[-300:] + r.stderr[-300:]).strip()}
    return {"value": float(m.group(1)), "unit": "zone-updates/s", "steps": nsteps, "process_wall_s": round(wall, 1),
            "what": "unmodified reference driver + shim + libathena_b200, its own main-loop zone-cycles/wall-second"}


def run_port_cpu(workload, n, nsteps):
    """Fallback when oracle/_ref is absent: the oracle port (oracle/athena_oracle.c), 1 thread."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    from athena_b200 import problems
    par = make_par(workload, n, n)
    nx = (n, n, n)
    xmin = tuple(par.par_getd("domain1", f"x{a}min") for a in (1, 2, 3))
    xmax = tuple(par.par_getd("domain1", f"x{a}max") for a in (1, 2, 3))
    bc = tuple(par.par_geti("domain1", k) for k in ("bc_ix1", "bc_ox1", "bc_ix2", "bc_ox2", "bc_ix3", "bc_ox3"))
    prob = WORKLOADS[workload][1]
    ion = prob != "blast"
    grav = None
    if prob == "ioniz_sphere":
        problems.ioniz_sphere_setup(par)
        grav = problems.lib().ab2_planet_pot
    G = ol.OracleGrid(nx, xmin, xmax, par.par_getd("problem", "gamma"), 1 if ion else 0, bc, grav=grav)
    U = problems.problem(prob, par, nx, xmin, G.dx)
    for v in G.U:
        G.U[v][...] = U[v]
    G.bvals()
    flux = par.par_getd("problem", "flux") if ion else 0.0
    if ion:
        G.bvals_ionrad(flux)
    cour, tlim = par.par_getd("time", "cour_no"), par.par_getd("time", "tlim")
    G.new_dt(cour, tlim)
    t0 = time.time()
    ol.run_main_loop(G, nsteps, cour, tlim, ol.ion_par_from(par, G.dx) if ion else None, flux)
    wall = time.time() - t0
    return n ** 3 * nsteps / wall, wall, "port"


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def run_reference_window(workload, n, warmup, steps, instances):
    """The reference binary over the SAME step window as the GPU arm (steps warmup .. warmup+steps from t=0): two batches of
    `instances` concurrent serial runs, nlim = warmup and nlim = warmup+steps; the window's cost is the difference of their
    main-loop wall times (main.c:713 'elapsed wall time', set-up excluded by construction).  Returns (zone-updates/s summed
    over instances, wall seconds of both batches, kind)."""
    exe = reference_exe(workload)
    deck = os.path.join(DECKS, WORKLOADS[workload][0])
    if not os.path.exists(exe):
        v, wall, kind = run_port_cpu(workload, min(n, 32), max(2, min(steps, 3)))
        return v, wall, kind

    def batch(nlim):
        procs, tmps = [], []
        for _ in range(instances):
            tmp = tempfile.mkdtemp(prefix="ath_ref_")
            tmps.append(tmp)
            cmd = [exe, "-i", deck, "job/maxout=0", "job/num_domains=1", f"domain1/Nx1={n}", f"domain1/Nx2={n}", f"domain1/Nx3={n}", f"time/nlim={nlim}"]
            procs.append(subprocess.Popen(cmd, cwd=tmp, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True))
        walls = []
        for p in procs:
            out, _ = p.communicate()
            m = re.search(r"elapsed wall time = (\S+) sec", out)
            if m:
                walls.append(float(m.group(1)))
        for tmp in tmps:
            subprocess.run(["rm", "-rf", tmp])
        return walls

    t0 = time.time()
    wa = batch(warmup) if warmup > 0 else [0.0] * instances
    wb = batch(warmup + steps)
    wall = time.time() - t0
    if not wb or len(wa) != len(wb):
        return 0.0, wall, "reference"
    # instance by instance (same order of launch is not the same core, so use the means)
    dt = sum(wb) / len(wb) - sum(wa) / len(wa)
    if dt <= 0:                                  # tiny samples: timer noise; fall back to the whole run from t=0
        return len(wb) * n ** 3 * (warmup + steps) / (sum(wb) / len(wb)), wall, "reference"
    return len(wb) * n ** 3 * steps / dt, wall, "reference"


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = min(os.cpu_count() or 1, args.cpu_instances) if args.cpu_instances > 0 else (os.cpu_count() or 1)
    n = args.cpu_n
    value, wall, kind = run_reference_window(args.workload, n, args.warmup, args.steps, cores)
    used = cores if kind == "reference" else 1
    line = {
        "impl": "reference", "metric": "zone-updates/s per hydro+ion-RT step", "value": value, "unit": "zone-updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * n ** 3 * used / value if value else None,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{WORKLOADS[args.workload][2]}; CPU sample {n}^3 per instance, steps {args.warmup}..{args.warmup + args.steps} from t=0 "
                               "(the GPU arm's step window; its grid is 256^3 per GPU)", "cpu": cpu_model()},
        "cpu_baseline": {"value": value, "unit": "zone-updates/s", "cores": used, "kind": kind, "cpu": cpu_model(),
                         "sample": f"{used} concurrent serial runs of the reference binary (no MPI in this image; ideal-scaling proxy for an MPI run), "
                                   f"{n}^3 zones each, steps {args.warmup}..{args.warmup + args.steps}: main-loop wall time of an nlim={args.warmup + args.steps} "
                                   f"batch minus that of an nlim={args.warmup} batch; {wall:.0f} s in all"},
        "e2e": {"value": value, "unit": "zone-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=OUT, flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------------
# parity gate: before anything is timed, the very path that is about to be timed (same library, same driver, same number
# of GPUs, NCCL exchange and reductions included) must reproduce the REFERENCE's dump of small cases bit for bit
# ---------------------------------------------------------------------------------------------------------
GATE_CASES = {"ioniz_sphere": ["ioniz64s_k2", "blast64s_k4"], "blast": ["blast64s_k4"], "ifront": ["ifront64s_k3"],
              "ioniz_smr": ["ioniz64s_k2"]}


def parity_gate(workload, rank, world, local, dist, riemann=0, order=2, pair=False):
    """Runs tests/golden/<case>.json (SHA-256 fingerprints of the reference binary's own dump: 32x32x64 zones, so each of up
    to 8 x3 slabs keeps 8 layers) through driver.Run on `world` GPUs, gathers the slabs and compares every variable,
    EdgeFlux, time, dt and the sub-cycle counts.  Returns the dict that goes into the JSON line; the caller exits non-zero
    unless bit_equal."""
    import numpy as np
    import torch
    from athena_b200.driver import Run
    from athena_b200.restart import fingerprint
    out = {"cases": [], "bit_equal": True, "against": "SHA-256 fingerprints of the reference binary's dump (tests/golden/*.json)", "n_gpus": world}
    if riemann != 0 or order != 2:
        return {"cases": [], "bit_equal": None, "note": "gate fixtures exist for the shipped configuration (PLM + Roe + H-correction) only"}
    # Grids as small as the gate's take the fused update+rates kernel by default; when the Grid about to be timed is large enough
    # for the rates / update pair with its sparse sub-cycle (csrc/ion.cu), the gate is run through that pair too
    force_pair = pair and "AB2_ION_FUSE" not in os.environ
    if force_pair:
        os.environ["AB2_ION_FUSE"] = "0"
    out["ion_path"] = "rates/update pair + sub-cycle memo" if (pair and os.environ.get("AB2_ION_FUSE") == "0" and os.environ.get("AB2_ION_SKIP") != "0") else "as chosen by the library for the gate's Grid size"
    try:
        _parity_cases(out, workload, world, local, dist)
    finally:
        if force_pair:
            os.environ.pop("AB2_ION_FUSE", None)
    return out


def _parity_cases(out, workload, world, local, dist):
    import numpy as np
    import torch
    from athena_b200.driver import Run
    from athena_b200.restart import fingerprint
    for name in GATE_CASES[workload]:
        gold = json.load(open(os.path.join(ROOT, "tests", "golden", name + ".json")))
        par = ParFile.par_open(os.path.join(DECKS, gold["deck"]))
        par.par_cmdline(gold["overrides"])
        R = Run(par, gold["build"], device=local, dist=dist)
        R.run(gold["nstep"])
        mine = R.G.download_active()
        names = [v for v in gold["levels"][0] if v != "edgeflux"]
        g = {v: mine[v] for v in names}
        ef = R.G.download_edgeflux() if "edgeflux" in gold["levels"][0] else None
        if world > 1:
            loc = torch.from_numpy(np.stack([mine[v] for v in names])).cuda()
            parts = [torch.empty_like(loc) for _ in range(world)]
            dist.all_gather(parts, loc)
            full = torch.cat(parts, dim=1).cpu().numpy()
            g = {v: full[n] for n, v in enumerate(names)}
            if ef is not None:      # ray rows of every slab, the last edge row from the top slab
                e = torch.from_numpy(ef).cuda()
                eparts = [torch.empty_like(e) for _ in range(world)]
                dist.all_gather(eparts, e)
                ef = torch.cat([x[:-1] for x in eparts[:-1]] + [eparts[-1]], dim=0).cpu().numpy()
        if ef is not None:
            g["edgeflux"] = ef
        fp = fingerprint(dict(time=R.time, dt=R.dt, nstep=R.nstep, grids=[g]), [r[0] for r in R.rad_log], vars_=tuple(gold["levels"][0]))
        bad = [v for v, ref in gold["levels"][0].items() if fp["levels"][0][v]["sha256"] != ref["sha256"]]
        for k in ("rad", "time", "dt", "nstep"):
            if fp[k] != gold[k]:
                bad.append(k)
        out["cases"].append({"case": name, "bit_equal": not bad, "differing": bad, "exchanges": R.comm.counters["exchanges"] if R.comm else 0})
        out["bit_equal"] = out["bit_equal"] and not bad
        R.G.close()
        if R.comm:
            R.comm.close()


# ---------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------
def bind_to_gpu_numa(local):
    """Run this rank on the CPUs nearest its GPU (NVML's affinity mask), so the pinned host buffers of the e2e leg are first
    touched on that socket: with 8 ranks each moving 1.8 GB per step through host memory, cross-socket copies are what limits
    the end-to-end scaling.  Best effort; returns the number of CPUs bound to, or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = ((os.cpu_count() or 64) + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * i + b for i, w in enumerate(mask) for b in range(64) if (int(w) >> b) & 1]
        allowed = set(os.sched_getaffinity(0))
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def gpu_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from athena_b200 import engine
    from athena_b200.driver import Run

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    numa_cpus = bind_to_gpu_numa(local) if world > 1 else None
    sampler = ClockSampler(local) if rank == 0 else None     # started first thing: nvidia-smi takes ~1 s to emit its first sample
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this engine has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    d = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
        d = dist
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    # the bench's own static potential (csrc/problems.c ab2_planet_pot) is re-entrant: tabulate it with several host threads
    os.environ.setdefault("AB2_GRAV_THREADS", str(max(1, min(32, (os.cpu_count() or 1) // max(1, world)))))
    zones_gate = args.n * args.n * (args.n // world if args.strong else args.n)
    parity = (parity_gate(args.workload, rank, world, local, d, args.riemann, args.order, pair=zones_gate >= (1 << 20))
              if not args.no_parity_gate else None)
    if parity is not None and parity["bit_equal"] is False:
        if rank == 0:
            print(json.dumps({"metric": "zone-updates/s per hydro+ion-RT step", "value": None, "n_gpus": world, "parity": parity,
                              "error": "parity gate failed: nothing was timed"}), file=OUT, flush=True)
        raise SystemExit(3)

    n = args.n
    nx3_global = n if args.strong else n * world         # --strong: SURVEY 8d cfg 5, the n^3 box cut into `world` x3 slabs
    if nx3_global % world:
        raise SystemExit(f"bench.py: --strong needs n divisible by the number of GPUs ({n} % {world})")
    par = make_par(args.workload, n, nx3_global)
    prob = WORKLOADS[args.workload][1]
    t_setup = time.time()
    if args.chunk_planes is not None:
        os.environ["AB2_CHUNK_PLANES"] = str(args.chunk_planes)
    R = Run(par, prob, device=local, dist=d, riemann=args.riemann, order=args.order)
    t_setup = time.time() - t_setup
    G = R.G
    zones_local = R.zones
    zones_total = zones_local * world
    nvar, grav = G.nvar, (prob == "ioniz_sphere")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        R.step()
    barrier()
    if sampler is not None:      # nvidia-smi needs ~1 s for its first sample: make sure it is emitting before the timed region
        t_wait = time.time()
        while not sampler.lines and time.time() - t_wait < 3.0:
            time.sleep(0.05)

    # ---- timed region: K steps, state resident in HBM ----
    G.profile_enable(True)
    G.profile_read(reset=True)
    l0 = G.kernel_launches
    it0 = len(R.rad_log)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    w0 = time.time()
    e0.record()
    for _ in range(args.steps):
        R.step()
    e1.record()
    barrier()
    w1 = time.time()
    ms = e0.elapsed_time(e1)
    prof = G.profile_read(reset=True)
    G.profile_enable(False)
    launches = G.kernel_launches - l0
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clocks = sampler.summary(w0, w1) if sampler else None
    n_iter = (sum(r[0] for r in R.rad_log[it0:]) / max(1, len(R.rad_log) - it0)) if R.ion else 0.0
    value = zones_total * args.steps / (ms * 1e-3)

    # ---- roofline of the dominant kernel ----
    peak, peak_src = measured_peaks()
    ktab = {}
    tot_ms = sum(v[0] for v in prof.values())
    for name, (kms, cnt) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
        bpz = kernel_bytes_per_zone(name, nvar, grav)
        ent = {"ms_total": round(kms, 4), "launches": cnt, "share": round(kms / tot_ms, 4) if tot_ms else None}
        if bpz:
            ent["alg_bytes_per_launch"] = bpz * zones_local
            ent["gbs"] = round(bpz * zones_local / (kms / cnt * 1e-3) / 1e9, 1)
        ktab[name] = ent
    # dominant kernel = largest total time among the kernels whose compulsory bytes per launch are known (the [memo] launches of
    # the sparse ion sub-cycle evaluate a data-dependent fraction of the grid and carry no fixed figure)
    dom = next((k for k in ktab if "gbs" in ktab[k]), None)
    roof = None
    if dom:
        prof_all = {}
        try:
            prof_all = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        except Exception:
            pass
        ent = prof_all.get(dom, {})
        # which ceiling binds comes from the committed ncu --set full row of the same kernel (profiles/*_ncu_full_table.md):
        # DRAM throughput, FP64 pipe and issue-slot utilisation -- the largest names the bound
        utils = {"hbm": ent.get("dram_pct"), "fp64": ent.get("fp64_pipe_pct"), "issue": ent.get("issue_pct")}
        known = {k: v for k, v in utils.items() if v is not None}
        bound = max(known, key=known.get) if known else "hbm"
        roof = {"bound": bound, "kernel": dom, "achieved": ktab[dom]["gbs"], "peak": peak, "unit": "GB/s",
                "frac": round(ktab[dom]["gbs"] / peak, 4), "traffic": ent.get("dram_bytes_per_launch"), "peak_source": peak_src,
                "alg_bytes_per_launch": ktab[dom]["alg_bytes_per_launch"], "avg_launch_ms": round(ktab[dom]["ms_total"] / ktab[dom]["launches"], 4),
                "share_of_kernel_time": ktab[dom]["share"],
                "dram_util": round(ent["dram_pct"] / 100.0, 4) if "dram_pct" in ent else None,
                "fp64_util": round(ent["fp64_pipe_pct"] / 100.0, 4) if "fp64_pipe_pct" in ent else None,
                "issue_util": round(ent["issue_pct"] / 100.0, 4) if "issue_pct" in ent else None,
                "note": "frac is on the kernel's own compulsory bytes (DESIGN.md section 4); `bound` is the largest of the kernel's ncu utilisations"}
        # the whole step at SURVEY 8(d)'s compulsory bytes per zone-update, per GPU
        b_alg = step_algorithmic_bytes(args.workload, n_iter)
        roof["step_alg_bytes_per_zone_update"] = b_alg
        roof["step_achieved"] = round(value / world * b_alg / 1e9, 1)
        roof["step_frac"] = round(value / world * b_alg / 1e9 / peak, 4)
        # one ion sub-cycle iteration = rates + update passes at 112 B per zone
        rk = "k_ion_rates" if "k_ion_rates" in ktab else "k_ion_rates[first]"      # the full rates pass ([first]: with the call's prologue folded in)
        if rk in ktab and "k_ion_update" in ktab:
            pair_ms = ktab[rk]["ms_total"] / ktab[rk]["launches"] + ktab["k_ion_update"]["ms_total"] / ktab["k_ion_update"]["launches"]
            roof["ion_iteration_ms"] = round(pair_ms, 4)
            roof["ion_iteration_frac_at_112B"] = round(112.0 * zones_local / (pair_ms * 1e-3) / 1e9 / peak, 4)
        if R.ion:
            # sparse sub-cycle (csrc/ion.cu): from the second iteration of a step on, chunks of inert cells re-use their records
            nskip, ntask, nchunk = G.ion_memo_stats()
            memo_k = [k for k in ktab if k.endswith("[memo]") or k == "k_ion_walk"]
            memo_n = max([ktab[k]["launches"] for k in memo_k if k != "k_ion_walk"], default=0)
            roof["ion_memo"] = {"chunks_reused_frac": round(nskip / max(1, nchunk), 4), "rays_evaluated_frac": round(ntask / max(1, R.nx[1] * R.nx[2]), 4),
                                "snapshot": "the last scan of the chunk records in the timed region", "sparse_iterations": memo_n,
                                "sparse_iteration_ms": round(sum(ktab[k]["ms_total"] for k in memo_k) / memo_n, 4) if memo_n else None,
                                "note": "step_frac and ion_iteration_frac_at_112B count SURVEY 8(d)'s bytes for every zone and iteration, "
                                        "re-used chunks included; ion_iteration_* is measured on the full first-iteration passes (whose rates pass, "
                                        "k_ion_rates[first], also carries the call's prologue: floors + save_energy_and_x)"}
        roof["per_kernel_ncu"] = {k: {"dram_util": round(v["dram_pct"] / 100.0, 3), "fp64_util": round(v["fp64_pipe_pct"] / 100.0, 3),
                                      "issue_util": round(v["issue_pct"] / 100.0, 3) if "issue_pct" in v else None}
                                  for k, v in prof_all.items() if k.startswith("k_") and "dram_pct" in v}

    # ---- e2e: host buffers in and out every step (the per-call coherence mode of the drop-in shim) ----
    e2e_steps = min(args.steps, args.e2e_steps)
    host = torch.empty(G.shape + (nvar,), dtype=torch.float64, pin_memory=True).numpy()
    host[...] = G.download_aos()
    barrier()
    e0.record()
    for _ in range(e2e_steps):
        G.upload_aos(host)
        R.step()
        engine.lib().ab2_download_cons_aos(G.h, host.ctypes.data_as(engine.dp))
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    e2e = {"value": zones_total * e2e_steps / (ms_e2e * 1e-3), "unit": "zone-updates/s",
           "h2d_bytes_per_step": int(host.nbytes), "d2h_bytes_per_step": int(host.nbytes), "steps": e2e_steps,
           "path": "ab2_upload_cons_aos(pinned host ConsS) -> step -> ab2_download_cons_aos, per step",
           "mode": "one Grid: copy in, step, copy out, strictly one after the other",
           "cpus_bound_near_gpu": numa_cpus}
    if world == 1 and args.e2e_members > 1:
        # One Grid cannot hide its own PCIe time (its step needs the whole upload, its download the whole step).  A host that
        # holds several independent Grids -- an ensemble, or the same deck at several parameter values -- can: each member
        # does exactly the per-step sequence above on its own stream from its own host thread (the C calls release the GIL),
        # so one member's copies overlap another's kernels.  Every step still uploads its input and downloads its result.
        e2e["single_grid_value"] = e2e["value"]
        members = [(R, host)]
        for _ in range(args.e2e_members - 1):
            Rm = Run(par, prob, device=local, dist=None, riemann=args.riemann, order=args.order)
            Rm.run(R.nstep)                                  # the same stretch of the run as the first member
            hm = torch.empty(G.shape + (nvar,), dtype=torch.float64, pin_memory=True).numpy()
            hm[...] = Rm.G.download_aos()
            members.append((Rm, hm))
        errs = []

        def member_loop(Rm, hm):
            try:
                torch.cuda.set_device(local)
                for _ in range(e2e_steps):
                    Rm.G.upload_aos(hm)
                    Rm.step()
                    engine.lib().ab2_download_cons_aos(Rm.G.h, hm.ctypes.data_as(engine.dp))
            except BaseException as e:       # noqa: BLE001 -- reported below, the bench line must still appear
                errs.append(repr(e))

        torch.cuda.synchronize()
        ths = [threading.Thread(target=member_loop, args=m) for m in members]
        ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tw = time.time()
        ee0.record()
        for th in ths:
            th.start()
        for th in ths:
            th.join()
        torch.cuda.synchronize()
        ee1.record()
        torch.cuda.synchronize()
        tw = time.time() - tw
        ms_ens = ee0.elapsed_time(ee1)
        if errs:
            e2e["ensemble_error"] = errs[0]
        else:
            e2e.update({"value": zones_total * e2e_steps * len(members) / (ms_ens * 1e-3), "steps": e2e_steps * len(members),
                        "mode": f"{len(members)} independent Grids of the workload in flight on one GPU, one host thread and one stream each; "
                                "every member step = upload its pinned ConsS array, step, download it (single_grid_value: one Grid alone)",
                        "members": len(members), "wall_ms": round(tw * 1e3, 2), "event_ms": round(ms_ens, 2)})
        for Rm, _ in members[1:]:
            Rm.G.close()

    # ---- CPU baseline beside it (rank 0, N=1): the reference's own binary on a bounded sample ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, wall, kind = run_reference_cpu(args.workload, args.cpu_n, args.cpu_steps, 1)
        cpu = {"value": v, "unit": "zone-updates/s", "cores": 1, "kind": kind,
               "sample": f"{args.cpu_n}^3 zones, {args.cpu_steps} steps from t=0, serial reference build (gcc -O3, no MPI in this image), {wall:.1f} s"}

    dropin = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        torch.cuda.synchronize()
        dropin = run_dropin(args.workload, n, max(200, args.warmup + args.steps))     # long enough to amortise its one-time set-up
        if dropin and "value" in dropin:
            dropin["what"] += ("; includes the first step's one-time work (potential tables from the host callback, Userwork discovery, first upload) "
                               "and the re-validations of the Userwork replay against the real host function at steps 1, 4, 16, 64, ... (each a full download)")

    if rank == 0:
        line = {
            "metric": "zone-updates/s per hydro+ion-RT step", "value": value, "unit": "zone-updates/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{WORKLOADS[args.workload][2]} {n}x{n}x{nx3_global}, steps {args.warmup}..{args.warmup + args.steps} from t=0",
                       "zones_per_gpu": zones_local, "parallelism": f"x3-slabs x{world}",
                       "variant": {"order": args.order, "riemann": args.riemann},
                       "ion_iterations_per_step": round(n_iter, 2), "l2": "working set >> L2 (state + scratch = "
                       f"{G.bytes_allocated / 2**30:.1f} GiB per GPU), no flush needed", "setup_s": round(t_setup, 1),
                       "device_bytes_per_gpu": int(G.bytes_allocated), "scratch_planes": G.scratch_planes,
                       "hydro_passes": "fused correct+final (k_cfinal)" if os.environ.get("AB2_HYDRO_FUSE", "0") == "1" else "k_correct3 + k_final pair",
                       "comm": (dict(R.comm.counters, where="NCCL inside libathena_b200 (csrc/comm.cu)") if R.comm else None)},
            "parity": parity,
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "dropin": dropin, "kernels": ktab,
        }
        print(json.dumps(line), file=OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def smr_parity_gate(rank, world, local, dist):
    """The SMR path about to be timed (SmrRun on one GPU, SmrRunSlabs over x3 cut planes on several) on the reference's own
    2-/3-level dumps (tests/golden/*smr*.npz, written by the reference's --enable-smr binary): every level, EdgeFlux, time,
    dt and the sub-cycle counts bit for bit."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from smr_cases import SmrCase
    from athena_b200.driver import SmrRun, SmrRunSlabs
    out = {"cases": [], "bit_equal": True, "against": "restart dumps of the reference's --enable-smr binary (tests/golden/*.npz)", "n_gpus": world}
    for name in ("blastsmr_k4", "ifrontsmr_k3", "ifrontsmr3_k3"):
        case = SmrCase(name)
        U0 = [lv["U0"] for lv in case.levels]
        try:
            R = SmrRunSlabs(case.par, case.build0, U0, dist, device=local) if world > 1 else SmrRun(case.par, case.build0, *U0, device=local)
        except ValueError as e:          # the cut planes of this many ranks cannot serve this small nest (smr_slab_layout's rules)
            out["cases"].append({"case": name, "bit_equal": None, "skipped": str(e)})
            continue
        R.run(case.nsteps)
        bad = []
        for n in range(case.nlevels):
            if world > 1:
                got = R.gather_level(n)
                ef = R.gather_level(n, "edgeflux") if R.ion else None
            else:
                got = R.grids[n].download_active()
                ef = R.grids[n].download_edgeflux() if R.ion else None
            for v in got:
                if not np.array_equal(got[v], case.levels[n]["U1"][v], equal_nan=True):
                    bad.append(f"L{n}:{v}")
            if ef is not None and not np.array_equal(ef[:-1], case.levels[n]["edgeflux1"][:-1]):
                bad.append(f"L{n}:edgeflux")
        if R.time != float(case.gold["time1"]) or R.dt != float(case.gold["dt1"]):
            bad.append("time/dt")
        if case.ion and [list(r[:3]) for r in R.rad_log] != [[int(x) for x in r[:3]] for r in case.gold["rad"]]:
            bad.append("rad")
        out["cases"].append({"case": name, "bit_equal": not bad, "differing": bad})
        out["bit_equal"] = out["bit_equal"] and not bad
    return out


def smr_arm(args):
    """SURVEY 8d cfg 4: the SHIPPED 5-level ioniz_sphere deck (80^3/52^3/96^3/176^3/320^3 = 39.8 M zones, static
    potential, Userwork reset, ion RT on every level), one GPU (driver.SmrRun) or every level cut on common x3 planes
    over N GPUs (driver.SmrRunSlabs, balanced cuts).  Strong scaling: the deck is fixed."""
    import torch
    import torch.distributed as dist
    from athena_b200.athinput import ParFile
    from athena_b200.driver import SmrRun, SmrRunSlabs
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    os.environ.setdefault("AB2_GRAV_THREADS", str(max(1, min(32, (os.cpu_count() or 1) // max(1, world)))))
    par = ParFile.par_open(os.path.join(DECKS, "ioniz_sphere_hires.athinput"))
    nlev = par.par_geti("job", "num_domains")
    zones_total = sum(par.par_geti(f"domain{n + 1}", "Nx1") * par.par_geti(f"domain{n + 1}", "Nx2") * par.par_geti(f"domain{n + 1}", "Nx3")
                      for n in range(nlev))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    parity = smr_parity_gate(rank, world, local, dist) if not args.no_parity_gate else None
    if parity is not None and parity["bit_equal"] is False:
        if rank == 0:
            print(json.dumps({"metric": "zone-updates/s per hydro+ion-RT step", "value": None, "n_gpus": world, "parity": parity,
                              "error": "parity gate failed: nothing was timed"}), file=OUT, flush=True)
        raise SystemExit(3)
    t_setup = time.time()
    if world > 1:
        R = SmrRunSlabs(par, "ioniz_sphere", None, dist, device=local, cuts="balanced")
        grids = [L["G"] for L in R.levels if L is not None]
        cuts = R.cuts
    else:
        R = SmrRun(par, "ioniz_sphere", None, device=local)
        grids, cuts = R.grids, [0, par.par_geti("domain1", "Nx3")]
    t_setup = time.time() - t_setup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None
    for _ in range(args.warmup):
        R.step()
    barrier()
    if sampler is not None:
        t_wait = time.time()
        while not sampler.lines and time.time() - t_wait < 3.0:
            time.sleep(0.05)
    l0 = sum(G.kernel_launches for G in grids)
    it0 = len(R.rad_log)
    for G in grids:
        G.profile_enable(True)
        G.profile_read(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    w0 = time.time()
    e0.record()
    for _ in range(args.steps):
        R.step()
    e1.record()
    barrier()
    w1 = time.time()
    ms = e0.elapsed_time(e1)
    launches = sum(G.kernel_launches for G in grids) - l0
    # per level: zones, sub-cycle iterations per step, kernel time per step (CUDA events inside the library), memo snapshot
    per_level = []
    nlev_here = len(grids)
    for n, G in enumerate(grids):
        prof = G.profile_read(reset=True)
        G.profile_enable(False)
        ion = {k: round(v[0] / args.steps, 3) for k, v in prof.items() if k.startswith("k_ion")}
        lv_its = [r[0] for r in R.rad_log[it0:]][n::nlev_here] if world == 1 else []
        nskip, ntask, nchunk = G.ion_memo_stats()
        per_level.append({"zones": G.nx[0] * G.nx[1] * G.nx[2], "iterations_per_step": round(sum(lv_its) / max(1, len(lv_its)), 2) if lv_its else None,
                          "ion_ms_per_step": ion,      # (the levels' hydro kernels overlap on their own streams: no per-level split)
                          "memo": {"chunks_reused_frac": round(nskip / max(1, nchunk), 4), "rays_listed_frac": round(ntask / max(1, G.nx[1] * G.nx[2]), 4)}})
    if world > 1:
        t = torch.tensor([ms, float(launches)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t[:1], op=dist.ReduceOp.MAX)
        dist.all_reduce(t[1:], op=dist.ReduceOp.SUM)
        ms, launches = float(t[0].item()), int(t[1].item())
    clocks = sampler.summary(w0, w1) if sampler else None
    its = [r[0] for r in R.rad_log[it0:]]
    peak, peak_src = measured_peaks()
    value = zones_total * args.steps / (ms * 1e-3)
    n_iter = sum(its) / max(1, len(its))
    b_alg = 128 + n_iter * 112                       # B_hydro (static potential) + n_iter * B_ion
    if rank == 0:
        line = {"metric": "zone-updates/s per hydro+ion-RT step", "value": value, "unit": "zone-updates/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"configs[3] ioniz_sphere SMR, shipped 5-level deck {zones_total} zones, steps {args.warmup}..{args.warmup + args.steps} from t=0",
                           "parallelism": f"x3 cut planes at root layers {cuts}", "ion_iterations_per_level_step": round(n_iter, 2),
                           "l2": "working set >> L2, no flush needed", "setup_s": round(t_setup, 1)},
                "per_level": per_level,
                "parity": parity, "e2e": None, "gpu_launches": int(launches), "clocks": clocks,
                "roofline": {"bound": "hbm", "kernel": "whole step", "achieved": round(value / world * b_alg / 1e9, 1), "peak": peak, "unit": "GB/s",
                             "frac": round(value / world * b_alg / 1e9 / peak, 4), "traffic": None, "peak_source": peak_src,
                             "note": "whole step at SURVEY 8(d) compulsory bytes per GPU; per-kernel rooflines are those of the unigrid workload"},
                "cpu_baseline": None}
        print(json.dumps(line), file=OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    # stdout carries exactly ONE line, the JSON record: whatever libraries print there (NCCL's version banner, torchrun
    # notices) goes to stderr instead
    global OUT
    OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="ioniz_sphere", choices=sorted(WORKLOADS) + ["ioniz_smr"])
    ap.add_argument("--n", "--size", dest="n", type=int, default=256, help="zones per direction per GPU (use --size under torchrun, whose own parser trips over --n)")
    ap.add_argument("--order", type=int, default=2, choices=[2, 3], help="reconstruction: 2 PLM (the shipped configuration), 3 PPM")
    ap.add_argument("--riemann", type=int, default=0, choices=[0, 1, 2], help="0 Roe + H-correction (shipped), 1 HLLC, 2 Roe without H-correction")
    ap.add_argument("--strong", action="store_true", help="--n is the GLOBAL cube size, cut into --gpus x3 slabs (strong scaling)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--e2e-members", type=int, default=3, help="independent Grids in flight in the e2e leg on one GPU (1: a single Grid, copies and step one after the other)")
    ap.add_argument("--chunk-planes", type=int, default=None, help="x3-chunked CTU scratch: active planes per chunk (default: whole Grid if it fits, else automatic)")
    ap.add_argument("--cpu-n", type=int, default=None, help="grid of the bounded CPU sample (default: 128 for --impl reference, 64 for the GPU arm's cpu_baseline)")
    ap.add_argument("--cpu-instances", type=int, default=0, help="--impl reference: concurrent serial instances (default: one per host core)")
    ap.add_argument("--cpu-steps", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-gate", action="store_true", help="skip the bit-exact check against the reference's fingerprints before the timed region")
    args = ap.parse_args()
    if args.cpu_n is None:
        args.cpu_n = 128 if args.impl == "reference" else 64
    if args.warmup < 3 and args.impl == "b200":
        print("bench.py: note: fewer than 3 warm-up steps", file=sys.stderr)
    if args.workload == "ioniz_smr":
        if args.impl == "reference":
            raise SystemExit("bench.py: the reference arm of the SMR deck is tests/golden/make_hires5_fingerprint.py (6 min per step on one core)")
        return smr_arm(args)
    return reference_arm(args) if args.impl == "reference" else gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
