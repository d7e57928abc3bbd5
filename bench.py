#!/usr/bin/env python
"""bench.py -- the headline benchmark of the PFEMFort implicit hot path on B200.

Default workload = BASELINE.json configs[4] (the configuration `metric` is quoted on): 3-D Poisson on the genTetra
200x200x200x6 = 48 M P1-tet mesh over [-1,1]^3, u = x^2+y^2+z^2 on the boundary, source -6, Jacobi-CG to rtol 1e-10
(reference run with -ksp_type cg -pc_type jacobi -ksp_rtol 1e-10).  `--workload c2|c3|c4|c5` selects the other
BASELINE configurations (c2 tria1000x1000 Poisson, c3 beam3Dtet6366 elasticity fixture with its ForceBC file,
c4 beam 50x300x50x6 elasticity).  Synthetic inputs are generated on the box by the reference's own recipes
(pfemfort_b200/mesh.py); c3 reads the bundled fixture under tests/golden/input.

One "step" = one pass of the hot path: setZero -> fused value pass (Ke/Fe + assembly + lifting) [-> ForceBC adds]
-> Jacobi-CG solve to tolerance.  `value` (CG DOF-iter/s = N_free * iterations / solve seconds, whole job over all
ranks) is measured with the mesh, pattern and applied values already resident in HBM; `assembly` carries the second
half of the metric (Melem/s).  `e2e` is the same metric through the C ABI from HOST buffers: mesh upload, pattern
pass, value pass, solve and the read-back of the solution are all inside its timed region.  `parity` compares the
GPU's assembled values / RHS / iteration count with the CPU oracle in the same run (N = 1: on the benchmark's own
system; N > 1: per-rank row blocks of a mid-size mesh, untimed, before the timed region).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c5] [--cells n]

For N > 1 launch with torchrun (one rank per GPU); rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

RTOL = 1e-10
# what holds the oracle itself to the reference
ORACLE_PIN = ("golden vectors from EXECUTING the reference's own Fortran (oracle/refrun; tests/golden/ref_*.npz): element routines, "
              "numbering, pattern, assembled matrix / RHS bit for bit; PETSc's Krylov iterates are not pinned by a reference run")
TUNING_ENV = ("PFEM_ASM", "PFEM_CG", "PFEM_CG_SR", "PFEM_PCG_CFG", "PFEM_PCG_FUSED", "PFEM_KERNELS_P2P", "PFEM_TILE_ROWS",
              "PFEM_TILE_THREADS", "PFEM_SYNC", "PFEM_ARITH", "PFEM_PCG_SYNC", "PFEM_PCG_HALO")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c5", choices=["c2", "c3", "c4", "c5"])
    ap.add_argument("--cells", type=int, default=0, help="size override: cells per side (c5: 200, c4: 50, c2: 1000)")
    ap.add_argument("--partition", default="metis", choices=["metis", "slab"])
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--cpu-elems", type=int, default=0, help="elements in the CPU assembly sample (0 = all)")
    ap.add_argument("--cpu-its", type=int, default=-1,
                    help="CG iterations in the CPU solve sample (-1 = workload default; 0 = solve to tolerance)")
    return ap.parse_args()


def host_threads() -> int:
    """Host threads of the CPU arm: every core this process may run on, regardless of OMP_NUM_THREADS
    (torchrun exports OMP_NUM_THREADS=1, which must not shrink the reference arm)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def algorithmic_bytes(npe, nsize, ndim, nElem, nNode, N, nnz, nDBC):
    """compulsory traffic, int32 indices, FP64 values."""
    asm = 4 * npe * nElem + 4 * nsize * nElem + 8 * ndim * nNode + 12 * nnz + 4 * N + 8 * N + 8 * nDBC
    spmv = 12 * nnz + 4 * (N + 1) + 16 * N
    cg_iter = 12 * nnz + 4 * (N + 1) + 104 * N
    return asm, spmv, cg_iter


def captured_traffic(kernel: str, workload: str, n_gpus: int):
    """dram__bytes_read+write per launch of `kernel` from a committed `ncu --set full` capture (profiles/traffic.json,
    written by tools/ncu_traffic.py from the .ncu-rep), or None: never a number that does not belong to this build /
    workload / tuning."""
    if any(os.environ.get(k) for k in TUNING_ENV):
        return None, None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tab = json.load(f)
        for rec in tab:
            if rec["workload"] == workload and rec["n_gpus"] == n_gpus and rec["kernel"] in kernel:
                return rec, rec.get("source")
    except Exception:
        pass
    return None, None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


# ---- workloads (BASELINE.json configs[1..4]) ----------------------------------------------------------------------

def make_workload(name: str, cells: int):
    """-> dict(mesh, kind, label, max_it, cpu_its): the inputs of one BASELINE configuration, by the reference's recipes."""
    from pfemfort_b200 import mesh as M, solver as S
    if name == "c5":
        n = cells or 200
        m = M.gen_tetra(-1.0, 1.0, n, -1.0, 1.0, n, -1.0, 1.0, n)
        return dict(mesh=m, kind=S.POISSON_TETRA, max_it=100000, cpu_its=0, n=n,
                    label=f"genTetra {n}^3 x6 P1-tet Poisson on [-1,1]^3, Jacobi-CG rtol {RTOL:g}")
    if name == "c4":
        n = cells or 50
        m = M.gen_tetra(-0.5, 0.5, n, 0.0, 6.0, 6 * n, -0.5, 0.5, n, dbc="clamp_y0", ndof=3)
        return dict(mesh=m, kind=S.ELASTICITY_TETRA, max_it=200000, cpu_its=200, n=n,
                    label=f"genTetra beam {n}x{6 * n}x{n} x6 P1-tet linear elasticity, clamped at y=0, Jacobi-CG rtol {RTOL:g}")
    if name == "c3":
        m = M.read_mesh(os.path.join(ROOT, "tests", "golden", "input", "beam3Dtet6366"), swap_34=True)
        return dict(mesh=m, kind=S.ELASTICITY_TETRA, max_it=100000, cpu_its=0, n=0,
                    label=f"beam3Dtet6366 fixture (local nodes 3<->4 swapped) P1-tet linear elasticity + ForceBC, Jacobi-CG rtol {RTOL:g}")
    n = cells or 1000
    m = M.gen_tria_poisson(n)
    return dict(mesh=m, kind=S.POISSON_TRIA, max_it=100000, cpu_its=0, n=n,
                label=f"tria{n}x{n} P1 Poisson on the unit square, Jacobi-CG rtol {RTOL:g}")


def node_partition(w, nparts, how, rank, bcast):
    from pfemfort_b200 import driver as D
    m = w["mesh"]
    if nparts <= 1:
        return None
    if how == "slab" and w["n"]:
        # slabs of whole node planes (rows in 2-D) along the slowest axis: a valid node partition (partition vectors
        # are inputs to the hot path)
        planes = w["n"] + 1
        nn = m.nNode // planes
        k = np.arange(m.nNode, dtype=np.int64) // nn
        return (k * nparts // planes).astype(np.int32)
    npart = np.zeros(m.nNode, np.int32)
    if rank == 0:      # rank 0 partitions, everyone receives (tetrapoissonparallelimpl1.F:457-484)
        _, npart = D.partition(m, w["kind"], nparts)
    return bcast(np.ascontiguousarray(npart, np.int32))


def force_bc(w, num):
    from pfemfort_b200 import driver as D, solver as S
    m = w["mesh"]
    if not m.fbc_node.size:
        return [], []
    return D.force_bc_rows(m, num, S.KIND_DIMS[w["kind"]][1])


def diff_stats(a, b):
    """max|a-b| relative to max|b| (norm-wise), and the count of bitwise-different entries"""
    if a.size == 0:
        return 0.0, 0
    scale = float(np.abs(b).max()) or 1.0
    d = 0.0
    nd = 0
    step = 1 << 24
    for i in range(0, a.size, step):
        x, y = a[i:i + step], b[i:i + step]
        d = max(d, float(np.abs(x - y).max()))
        nd += int(np.count_nonzero(x != y))
    return d / scale, nd


def oracle_system(w, num, threads, elem_mask=None, val=None, rhs=None, rp=None, col=None):
    """The CPU oracle's assembled system for a workload (pattern unless given, value pass, ForceBC adds)."""
    from oracle import pyoracle as O
    from pfemfort_b200 import driver as D
    m, kind = w["mesh"], w["kind"]
    if rp is None:
        rp, col = O.pattern(num.elemDof, num.size_global)
    old = num.node_map_get_old if num.nparts > 1 else None
    val, rhs, nbad = O.assemble(kind, num.conn_new, m.coords, old, num.elemDof, num.solnApplied, D.DEFAULT_ELEMDATA[kind],
                                D.DEFAULT_TIMEDATA, rp, col, elem_mask=elem_mask, threads=threads, val=val, rhs=rhs)
    rows, vals = force_bc(w, num)
    for r, v in zip(rows, vals):
        rhs[r] += v
    return rp, col, val, rhs


def run_reference(args, rank, world):
    """The reference arm: the CPU restatement of the PETSc/MPI path (oracle/pfem_oracle.c, OpenMP) on ALL host cores
    of the box.  PETSc, MPI and a Fortran compiler are absent here and on the GPU box, so oracle/_ref cannot exist;
    kind = "port".  Each step = a bounded sample: value pass over the first `cpu_elems` elements + `cpu_its` CG
    iterations on the full system.  Under torchrun only rank 0 works."""
    if rank != 0:
        return
    from oracle import pyoracle as O
    from pfemfort_b200 import driver as D
    threads = host_threads()
    w = make_workload(args.workload, args.cells)
    m, kind = w["mesh"], w["kind"]
    num = D.number(m, kind)
    N = num.size_global
    # full assembly once (untimed set-up of the CG sample's matrix), threaded
    rp, col, val, rhs = oracle_system(w, num, threads)
    ne = min(args.cpu_elems or m.nElem, m.nElem)
    its_cap = args.cpu_its if args.cpu_its >= 0 else (100 if args.workload in ("c5", "c4") else 0)
    mask = np.zeros(m.nElem, np.uint8)
    mask[:ne] = 1
    t_asm, t_cg, its_done = [], [], 0
    for step in range(args.warmup + args.steps):
        v2 = np.zeros_like(val)
        r2 = np.zeros_like(rhs)
        t0 = time.perf_counter()
        O.assemble(kind, num.conn_new, m.coords, None, num.elemDof, num.solnApplied, D.DEFAULT_ELEMDATA[kind],
                   D.DEFAULT_TIMEDATA, rp, col, elem_mask=mask, threads=threads, val=v2, rhs=r2)
        t1 = time.perf_counter()
        _, its_done, _, _ = O.cg_jacobi(rp, col, val, rhs, rtol=RTOL, max_it=w["max_it"], threads=threads, fixed_its=its_cap)
        t2 = time.perf_counter()
        if step >= args.warmup:
            t_asm.append(t1 - t0)
            t_cg.append(t2 - t1)
    its_done = its_cap if its_cap else its_done
    cg_rate = N * its_done / float(np.mean(t_cg))
    asm_rate = ne / float(np.mean(t_asm)) / 1e6
    sample = (f"first {ne} of {m.nElem} elements assembled; {its_done} CG iterations on the full {N}-DOF system"
              + ("" if its_cap else " (solve to tolerance)"))
    line = {
        "impl": "reference", "metric": "poisson_cg_dof_iter_per_s", "value": cg_rate, "unit": "DOF-iter/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * float(np.mean(t_asm) + np.mean(t_cg)), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["label"], "elements": int(m.nElem), "nodes": int(m.nNode), "dof": int(N), "nnz": int(col.size)},
        "run": {"timing": "host wall clock (CPU arm)", "threads": threads},
        "assembly": {"metric": "assembly_melem_per_s", "value": asm_rate, "unit": "Melem/s"},
        "cpu_baseline": {"value": cg_rate, "unit": "DOF-iter/s", "cores": threads, "kind": "port", "sample": sample,
                         "assembly_melem_per_s": asm_rate,
                         "note": "CPU restatement of the PETSc/MPI path (PETSc, MPI, gfortran unavailable): OpenMP CSR Jacobi-CG + element loop with sorted-row insert"},
        "e2e": {"value": cg_rate, "unit": "DOF-iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def multirank_parity(args, rank, world, local_rank, new_id, bcast, gather_max, gather_min):
    """N > 1: untimed per-rank row-block check on a mid-size mesh of the same kind, against the sequential oracle.
    Every rank drives its GPU through the C ABI exactly like the timed run (same partitioner, same exchange path)."""
    from oracle import pyoracle as O
    from pfemfort_b200 import driver as D, solver as S
    small = {"c5": ("c5", 24), "c4": ("c4", 6), "c3": ("c3", 0), "c2": ("c2", 96)}[args.workload]
    w = make_workload(*small)
    m, kind = w["mesh"], w["kind"]
    npart = node_partition(w, world, args.partition, rank, bcast)
    num = D.number(m, kind, world, npart)
    s = S.SolverB200(device=local_rank, rank=rank, nranks=world, nccl_id=new_id())
    info = D.run_rank(s, m, num, rank=rank, rtol=RTOL, max_it=w["max_it"])
    rp, col, val = s.get_csr()
    rhs = s.get_rhs()
    x = s.get_solution()
    mode = s.assembly_mode()[0]
    s.free()
    lo, hi = num.row_range(rank)
    orp, ocol, oval, orhs = oracle_system(w, num, 1)
    ox, oits, oreason, _ = O.cg_jacobi(orp, ocol, oval, orhs, rtol=RTOL, max_it=w["max_it"])
    a, b = int(orp[lo]), int(orp[hi])
    pat_ok = bool(np.array_equal(rp, orp[lo:hi + 1] - orp[lo]) and np.array_equal(col, ocol[a:b]))
    scale = float(np.abs(oval).max()) or 1.0
    dv = float(np.abs(val - oval[a:b]).max()) / scale if pat_ok and b > a else (0.0 if pat_ok else float("inf"))
    dr = float(np.abs(rhs - orhs[lo:hi]).max()) / (float(np.abs(orhs).max()) or 1.0) if hi > lo else 0.0
    dx = float(np.abs(x - ox).max()) / (float(np.abs(ox).max()) or 1.0)
    bit = bool(pat_ok and np.array_equal(val, oval[a:b]) and np.array_equal(rhs, orhs[lo:hi]))
    return {"mesh": w["label"], "ranks": world, "pattern_equal_all_ranks": bool(gather_min(1.0 if pat_ok else 0.0) > 0.5),
            "values_max_rel": gather_max(dv), "rhs_max_rel": gather_max(dr), "solution_max_rel": gather_max(dx),
            "values_bit_identical_all_ranks": bool(gather_min(1.0 if bit else 0.0) > 0.5),
            "its_gpu": int(info["its"]), "its_oracle": int(oits), "reason_gpu": int(info["reason"]), "reason_oracle": int(oreason),
            "assembly_mode": mode, "oracle": "sequential (1 thread, reference element order)"}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from pfemfort_b200 import driver as D, solver as S

    S.load_library()
    if not torch.cuda.is_available() or S.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; libpfemb200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend="cpu:gloo,cuda:nccl", rank=rank, world_size=world)

    def new_id():
        if world == 1:
            return None
        idt = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            idt = torch.frombuffer(bytearray(S.comm_unique_id()), dtype=torch.uint8).clone()
        dist.broadcast(idt, 0)
        return bytes(idt.numpy().tobytes())

    def bcast(arr):
        if world == 1:
            return arr
        t = torch.from_numpy(arr)
        dist.broadcast(t, 0)
        return t.numpy()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def reduce_ranks(v: float, op) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return float(t.item())

    max_over_ranks = lambda v: reduce_ranks(v, dist.ReduceOp.MAX) if world > 1 else v
    min_over_ranks = lambda v: reduce_ranks(v, dist.ReduceOp.MIN) if world > 1 else v
    sum_over_ranks = lambda v: reduce_ranks(v, dist.ReduceOp.SUM) if world > 1 else v

    parity = None
    if world > 1 and not args.no_parity:
        parity = multirank_parity(args, rank, world, local_rank, new_id, bcast, max_over_ranks, min_over_ranks)

    t_setup0 = time.perf_counter()
    w = make_workload(args.workload, args.cells)
    m, kind = w["mesh"], w["kind"]
    npe, ndof, ndim = S.KIND_DIMS[kind]
    npart = node_partition(w, world, args.partition, rank, bcast)
    num = D.number(m, kind, world, npart)
    ed, td = D.DEFAULT_ELEMDATA[kind], D.DEFAULT_TIMEDATA
    lo, hi = num.row_range(rank)
    size_local = hi - lo
    if world > 1:
        lst = D.local_elements(num, rank)
        conn = np.ascontiguousarray(num.conn_new[:, lst])
        node_map = num.node_map_get_old
    else:
        conn, node_map = num.conn_new, None
    fbc_rows, fbc_vals = force_bc(w, num)
    # pinned host staging of the step inputs (the e2e leg copies from these every step)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    conn_p, coords_p, applied_p = pin(conn), pin(m.coords), pin(num.solnApplied)
    map_p = pin(node_map) if node_map is not None else None
    nda_p = pin(num.NodeDofArrayNew)
    xout_p = torch.empty(num.size_global, dtype=torch.float64).pin_memory().numpy()
    h2d_bytes = conn_p.nbytes + nda_p.nbytes + coords_p.nbytes + applied_p.nbytes + (map_p.nbytes if map_p is not None else 0)
    d2h_bytes = xout_p.nbytes

    s = S.SolverB200(device=local_rank, rank=rank, nranks=world, nccl_id=new_id())

    stage_t = {}

    def timed(name, fn, *a):
        t = time.perf_counter()
        r = fn(*a)
        stage_t[name] = stage_t.get(name, 0.0) + time.perf_counter() - t
        return r

    def upload_and_pattern():
        s.initialise(size_local, num.size_global)
        s.set_options(rtol=RTOL, max_it=w["max_it"], pc_type=S.PC_JACOBI)
        timed("set_mesh", s.set_mesh, kind, conn_p, coords_p, map_p)
        timed("set_pattern", s.set_pattern_nodal, nda_p)     # element dof lists are formed on the GPU
        timed("set_applied", s.set_applied, applied_p)

    def hot_step():
        s.setZero()
        s.assemble(ed, td)
        for r_, v_ in zip(fbc_rows, fbc_vals):               # rows outside this rank's block are skipped by the library
            s.add_value(r_, v_)
        s.factoriseAndSolve()
        return s.info()

    upload_and_pattern()
    t_setup = time.perf_counter() - t_setup0
    # ---- resident-input timing: W warm-up steps, then exactly K timed steps ----
    for _ in range(args.warmup):
        info = hot_step()
    s.launch_count(reset=True)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    t0 = time.perf_counter()
    t_asm = t_solve = 0.0
    its = 0
    for _ in range(args.steps):
        info = hot_step()
        t_asm += info["t_assemble"]
        t_solve += info["t_solve"]
        its += info["its"]
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    launches = s.launch_count()
    # stand-alone SpMV probe (event-timed launches of the SELL kernel alone, after the timed region)
    spmv_avg = s.time_spmv(20)
    wall = max_over_ranks(wall)
    t_asm = max_over_ranks(t_asm)
    t_solve = max_over_ranks(t_solve)
    launches_total = int(sum_over_ranks(float(launches)))
    N = num.size_global
    its_per_step = its / args.steps
    value = N * its / t_solve
    asm_value = m.nElem * args.steps / t_asm / 1e6

    # ---- sizes for the roofline (this rank's share) ----
    import ctypes as C
    nnz = C.c_longlong()
    S._chk(s._lib.pfem_solver_get_nnz(s._h, C.byref(nnz)))
    nnz_local = nnz.value
    nnz_total = int(sum_over_ranks(float(nnz_local)))
    peak, peak_src = measured_peak()
    asm_b, spmv_b, cgit_b = algorithmic_bytes(npe, npe * ndof, ndim, conn.shape[1], m.nNode, size_local, nnz_local, m.dbc_node.size)
    spmv_gbs = spmv_b / spmv_avg / 1e9 if spmv_avg > 0 else 0.0
    asm_gbs = asm_b * args.steps / (t_asm if t_asm > 0 else 1) / 1e9
    cgit_gbs = cgit_b * its / t_solve / 1e9

    # FP64 pipe of the value pass (BASELINE.json north_star: "with the FP64 pipe reported for the element kernels"):
    # FP64 instructions per (row, element) visit of the kernel that ran, as the library reports for its own build: the
    # DYNAMIC count (ncu sm__inst_executed_pipe_fp64 / visits; the static SASS count of the loop body, which includes the rarely
    # taken lifting branches, is in profiles/r02_sass_hot_kernels.txt), peak = SMs x 64 FP64 lanes x the SM clock under load.
    asm_fp64 = None
    asm_kernel = None
    try:
        am = s.assembly_info()
        asm_kernel = am["kernel"]
        sm_count = torch.cuda.get_device_properties(local_rank).multi_processor_count
        mhz = float((clocks or {}).get("sm_mhz") or (clocks or {}).get("sm_max_mhz") or 1965.0)
        fp64_ops = am["fp64_per_visit"] * am["visits"]
        fp64_peak = sm_count * 64 * mhz * 1e6
        fp64_rate = fp64_ops * args.steps / (t_asm if t_asm > 0 else 1)
        asm_fp64 = {"ops_per_launch": fp64_ops, "achieved_gops": fp64_rate / 1e9, "peak_gops": fp64_peak / 1e9,
                    "frac": fp64_rate / fp64_peak, "unit": "FP64 instr/s", "visits_per_element": am["visits"] / max(conn.shape[1], 1),
                    "fp64_instr_per_visit": am["fp64_per_visit"], "arith": am["arith"],
                    "peak_source": f"{sm_count} SMs x 64 lanes x {mhz:.0f} MHz (SM clock sampled under load)"}
    except Exception:        # reporting only: never fail the bench on it
        asm_fp64 = None

    # ---- end-to-end leg: host buffers in, solution out, every step ----
    e2e_t = 0.0
    e2e_its = 0
    e2e_steps = max(1, args.e2e_steps)
    launches_e2e = 0
    for k in range(e2e_steps + 1):          # first pass is the warm-up of this leg
        barrier()
        s.launch_count(reset=True)
        t0 = time.perf_counter()
        if k == 1:
            stage_t.clear()
        upload_and_pattern()
        info = timed("hot_step", hot_step)
        timed("get_solution", s.get_solution, xout_p)
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        if k > 0:
            e2e_t += dt
            e2e_its += info["its"]
            launches_e2e += s.launch_count()
    e2e_value = N * e2e_its / e2e_t

    # ---- solution sanity on every run ----
    u = D.nodal_solution(num, xout_p)
    if args.workload == "c5":
        err = float(np.abs(u[0] - (m.coords ** 2).sum(0)).max())          # nodally ~exact quadratic
    elif args.workload == "c2":
        from pfemfort_b200 import mesh as M
        err = float(np.abs(u[0] - M.exact_poisson_tria(m.coords[0], m.coords[1])).max())
    else:
        err = None
    max_u = float(np.sqrt((u ** 2).sum(0)).max())

    cg_kernel = "cg_persistent_kernel"
    trec, tsrc = captured_traffic(cg_kernel, args.workload, world)
    arec, asrc = captured_traffic(asm_kernel or "assemble", args.workload, world)
    line = {
        "metric": "poisson_cg_dof_iter_per_s", "value": value, "unit": "DOF-iter/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["label"], "elements": int(m.nElem), "nodes": int(m.nNode), "dof": int(N), "nnz": nnz_total},
        "run": {"partition": "none" if world == 1 else args.partition, "parallelism": f"rows{world}",
                "exchange": {0: "none", 1: "nccl", 2: "peer-memory kernels (NVLink)"}[s.comm_mode()],
                "l2": "inputs larger than L2 (matrix >= 1.4 GB per pass vs 126 MB L2); no flush needed" if nnz_local * 12 > 2.5e8
                      else "per-rank matrix below 2x L2: the solve is L2-assisted at this size (no flush between iterations of one solve)",
                "timing": "CUDA events on the library stream (t_assemble, t_solve), max over ranks; ms_per_step = host wall between barriers",
                "tuning_env": {k: os.environ[k] for k in TUNING_ENV if os.environ.get(k)}},
        "iterations_per_step": its_per_step, "reason": info["reason"], "max_nodal_error": err, "max_abs_u": max_u,
        "assembly": {"metric": "assembly_melem_per_s", "value": asm_value, "unit": "Melem/s",
                     "ms_per_pass": 1e3 * t_asm / args.steps,
                     "roofline": {"bound": "hbm", "achieved": asm_gbs, "peak": peak, "unit": "GB/s", "frac": asm_gbs / peak,
                                  "traffic": arec["bytes_per_launch"] if arec else None, "traffic_source": asrc,
                                  "bytes_per_launch": asm_b, "peak_source": peak_src, "scope": "rank 0 share"},
                     "kernel": asm_kernel, "fp64_pipe": asm_fp64},
        "cg_iteration": {"ms_per_iteration": 1e3 * t_solve / max(its, 1), "achieved_gbs": cgit_gbs, "frac": cgit_gbs / peak,
                         "bytes_per_iteration": cgit_b},
        # dominant kernel = the persistent CG kernel (one cooperative launch per solve: set-up + every iteration);
        # duration = CUDA events on the launching stream around that launch, live in the timed steps above
        "roofline": {"bound": "hbm", "kernel": cg_kernel, "achieved": cgit_gbs, "peak": peak, "unit": "GB/s",
                     "frac": cgit_gbs / peak,
                     "traffic": (trec["bytes_per_iteration"] * its_per_step if trec else None), "traffic_source": tsrc,
                     "bytes_per_launch": cgit_b * its_per_step,
                     "avg_launch_us": 1e6 * t_solve / args.steps, "launches_timed": args.steps,
                     "bytes_per_iteration": cgit_b, "peak_source": peak_src, "scope": "rank 0 share"},
        "spmv_probe": {"kernel": "spmv_sell_kernel (stand-alone launches)", "achieved_gbs": spmv_gbs, "frac": spmv_gbs / peak,
                       "bytes_per_launch": spmv_b, "avg_launch_us": 1e6 * spmv_avg, "launches_timed": 20},
        "e2e": {"value": e2e_value, "unit": "DOF-iter/s", "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(d2h_bytes),
                "ms_per_step": 1e3 * e2e_t / e2e_steps, "steps": e2e_steps,
                "includes": "mesh upload, pattern pass, value pass, solve, solution read-back",
                "stage_ms": {k_: 1e3 * v_ / e2e_steps for k_, v_ in stage_t.items()}},
        "gpu_launches": launches_total, "gpu_launches_e2e": launches_e2e,
        "setup_s": t_setup, "clocks": clocks,
    }
    if parity is not None:
        if isinstance(parity, dict):
            parity.setdefault("oracle_pinned_by", ORACLE_PIN)
        line["parity"] = parity

    # ---- CPU baseline beside it (rank 0, N = 1 only): bounded sample of the same workload, and the parity record ----
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        from oracle import pyoracle as O
        threads = host_threads()
        rp, col, val = s.get_csr()
        rhs = s.get_rhs()
        ne = min(args.cpu_elems or m.nElem, m.nElem)
        mask = np.zeros(m.nElem, np.uint8)
        mask[:ne] = 1
        v2, r2 = np.zeros_like(val), np.zeros_like(rhs)
        t0 = time.perf_counter()
        O.assemble(kind, num.conn_new, m.coords, None, num.elemDof, num.solnApplied, ed, td, rp, col, elem_mask=mask,
                   threads=threads, val=v2, rhs=r2)
        t1 = time.perf_counter()
        for r_, v_ in zip(fbc_rows, fbc_vals):
            r2[r_] += v_
        its_cap = args.cpu_its if args.cpu_its >= 0 else w["cpu_its"]
        _, oits, oreason, _ = O.cg_jacobi(rp, col, val, rhs, rtol=RTOL, max_it=w["max_it"], threads=threads, fixed_its=its_cap)
        t2 = time.perf_counter()
        its_done = its_cap if its_cap else oits
        line["cpu_baseline"] = {
            "value": N * its_done / (t2 - t1), "unit": "DOF-iter/s", "cores": threads, "kind": "port",
            "assembly_melem_per_s": ne / (t1 - t0) / 1e6,
            "sample": f"first {ne} of {m.nElem} elements assembled ({t1 - t0:.1f} s); {its_done} CG iterations on the full system ({t2 - t1:.1f} s)"
                      + ("" if its_cap else ", solve to tolerance"),
            "note": "CPU restatement of the PETSc/MPI path (PETSc unavailable): OpenMP oracle on the box's host cores"}
        if not args.no_parity:
            # the threaded oracle adds with atomics (unordered): agreement to rounding, not bit-identity, is the claim
            par = {"oracle": f"OpenMP element loop, {threads} threads (unordered adds: agreement to rounding is the claim)",
                   "its_gpu": int(info["its"]), "its_oracle": (int(oits) if not its_cap else None),
                   "reason_gpu": int(info["reason"]), "reason_oracle": (int(oreason) if not its_cap else None),
                   "tolerance": 1e-12, "oracle_pinned_by": ORACLE_PIN}
            if ne == m.nElem:
                dv, nv = diff_stats(val, v2)
                dr, nr = diff_stats(rhs, r2)
                par.update({"values_max_rel": dv, "rhs_max_rel": dr, "values_differing_entries": nv, "rhs_differing_entries": nr,
                            "within_tolerance": bool(dv <= 1e-12 and dr <= 1e-12)})
            line["parity"] = par
    s.free()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
