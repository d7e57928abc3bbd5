#!/usr/bin/env python
"""bench.py -- the headline benchmark of the PFEMFort implicit hot path on B200.

Workload (BASELINE.json configs[4], the configuration `metric` is quoted on): 3-D Poisson on the genTetra
200x200x200x6 = 48 M P1-tet mesh over [-1,1]^3, u = x^2+y^2+z^2 on the boundary, source -6, Jacobi-CG to
rtol 1e-10 (reference run with -ksp_type cg -pc_type jacobi -ksp_rtol 1e-10).  Synthetic inputs generated on
the box by the reference's own recipe (pfemfort_b200/mesh.py).

One "step" = one pass of the hot path: setZero -> fused value pass (Ke/Fe + assembly + lifting) -> Jacobi-CG
solve to tolerance.  `value` (CG DOF-iter/s = N_free * iterations / solve seconds, whole job over all ranks)
is measured with the mesh, pattern and applied values already resident in HBM; `assembly` carries the second
half of the metric (Melem/s).  `e2e` is the same metric through the C ABI from HOST buffers: mesh upload,
pattern pass, value pass, solve and the read-back of the solution are all inside its timed region.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--cells 200]

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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cells", type=int, default=200, help="cells per side of the genTetra cube (200 = 48 M tets)")
    ap.add_argument("--partition", default="metis", choices=["metis", "slab"])
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-elems", type=int, default=48_000_000, help="elements in the CPU assembly sample")
    ap.add_argument("--cpu-its", type=int, default=100, help="CG iterations in the CPU solve sample")
    return ap.parse_args()


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def algorithmic_bytes(nElem, nNode, N, nnz, nDBC):
    """compulsory traffic, int32 indices, FP64 values."""
    asm = 16 * nElem + 16 * nElem + 8 * 3 * nNode + 12 * nnz + 4 * N + 8 * N + 8 * nDBC
    spmv = 12 * nnz + 4 * (N + 1) + 16 * N
    cg_iter = 12 * nnz + 4 * (N + 1) + 104 * N
    return asm, spmv, cg_iter


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


def build_workload(cells, nparts, partition, rank, bcast):
    from pfemfort_b200 import driver as D, mesh as M, solver as S
    m = M.gen_tetra(-1.0, 1.0, cells, -1.0, 1.0, cells, -1.0, 1.0, cells)
    npart = None
    if nparts > 1:
        if partition == "slab":
            # plane-wise slabs in z: a valid node partition (partition vectors are inputs to the hot path)
            nn = (cells + 1) * (cells + 1)
            k = np.arange(m.nNode, dtype=np.int64) // nn
            npart = (k * nparts // (cells + 1)).astype(np.int32)
        else:
            npart = np.zeros(m.nNode, np.int32)
            if rank == 0:      # rank 0 partitions, everyone receives (tetrapoissonparallelimpl1.F:457-484)
                _, npart = D.partition(m, S.POISSON_TETRA, nparts)
            npart = bcast(npart)
    num = D.number(m, S.POISSON_TETRA, nparts, npart)
    return m, num


def run_reference(args, rank, world):
    """The reference arm: the CPU restatement of the PETSc/MPI path (oracle/pfem_oracle.c, OpenMP) on the host
    cores.  PETSc, MPI and a Fortran compiler are absent here and on the GPU box, so oracle/_ref cannot exist;
    kind = "port".  Each step = a bounded sample: value pass over the first `cpu_elems` elements + `cpu_its`
    CG iterations on the full system."""
    if rank != 0:
        return
    from oracle import pyoracle as O
    from pfemfort_b200 import driver as D, mesh as M, solver as S
    threads = O.num_threads()
    m = M.gen_tetra(-1.0, 1.0, args.cells, -1.0, 1.0, args.cells, -1.0, 1.0, args.cells)
    o = O.number_dofs(m.nNode, 1, m.dbc_node, m.dbc_dof, m.dbc_val)
    conn_new = m.conn
    edof = O.elem_dof_array(conn_new, o["NodeDofArrayNew"])
    N = o["size_global"]
    rp, col = O.pattern(edof, N)
    ed, td = D.DEFAULT_ELEMDATA[S.POISSON_TETRA], D.DEFAULT_TIMEDATA
    # full assembly once (untimed set-up of the CG sample's matrix), threaded
    val, rhs, _ = O.assemble(O.POISSON_TETRA, conn_new, m.coords, None, edof, o["solnApplied"], ed, td, rp, col, threads=threads)
    ne = min(args.cpu_elems, m.nElem)
    mask = np.zeros(m.nElem, np.uint8)
    mask[:ne] = 1
    t_asm, t_cg = [], []
    for step in range(args.warmup + args.steps):
        v2 = np.zeros_like(val)
        r2 = np.zeros_like(rhs)
        t0 = time.perf_counter()
        O.assemble(O.POISSON_TETRA, conn_new, m.coords, None, edof, o["solnApplied"], ed, td, rp, col, elem_mask=mask,
                   threads=threads, val=v2, rhs=r2)
        t1 = time.perf_counter()
        O.cg_jacobi(rp, col, val, rhs, rtol=RTOL, threads=threads, fixed_its=args.cpu_its)
        t2 = time.perf_counter()
        if step >= args.warmup:
            t_asm.append(t1 - t0)
            t_cg.append(t2 - t1)
    cg_rate = N * args.cpu_its / float(np.mean(t_cg))
    asm_rate = ne / float(np.mean(t_asm)) / 1e6
    sample = f"first {ne} of {m.nElem} elements assembled; {args.cpu_its} CG iterations on the full {N}-DOF system"
    line = {
        "impl": "reference", "metric": "poisson_cg_dof_iter_per_s", "value": cg_rate, "unit": "DOF-iter/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * float(np.mean(t_asm) + np.mean(t_cg)), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"genTetra {args.cells}^3 x6 P1-tet Poisson on [-1,1]^3, Jacobi-CG rtol {RTOL:g}",
                   "elements": int(m.nElem), "dof": int(N), "nnz": int(col.size), "timing": "host wall clock (CPU arm)"},
        "assembly": {"metric": "assembly_melem_per_s", "value": asm_rate, "unit": "Melem/s"},
        "cpu_baseline": {"value": cg_rate, "unit": "DOF-iter/s", "cores": threads, "kind": "port", "sample": sample,
                         "assembly_melem_per_s": asm_rate,
                         "note": "CPU restatement of the PETSc/MPI path (PETSc, MPI, gfortran unavailable): OpenMP CSR Jacobi-CG + element loop with sorted-row insert"},
        "e2e": {"value": cg_rate, "unit": "DOF-iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


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
    nccl_id = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend="cpu:gloo,cuda:nccl", rank=rank, world_size=world)
        idt = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            idt = torch.frombuffer(bytearray(S.comm_unique_id()), dtype=torch.uint8).clone()
        dist.broadcast(idt, 0)
        nccl_id = bytes(idt.numpy().tobytes())

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

    def max_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    t_setup0 = time.perf_counter()
    m, num = build_workload(args.cells, world, args.partition, rank, bcast)
    kind = S.POISSON_TETRA
    ed, td = D.DEFAULT_ELEMDATA[kind], D.DEFAULT_TIMEDATA
    lo, hi = num.row_range(rank)
    size_local = hi - lo
    if world > 1:
        lst = D.local_elements(num, rank)
        conn = np.ascontiguousarray(num.conn_new[:, lst])
        node_map = num.node_map_get_old
    else:
        conn, node_map = num.conn_new, None
    # pinned host staging of the step inputs (the e2e leg copies from these every step)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    conn_p, coords_p, applied_p = pin(conn), pin(m.coords), pin(num.solnApplied)
    map_p = pin(node_map) if node_map is not None else None
    nda_p = pin(num.NodeDofArrayNew)
    xout_p = torch.empty(num.size_global, dtype=torch.float64).pin_memory().numpy()
    h2d_bytes = conn_p.nbytes + nda_p.nbytes + coords_p.nbytes + applied_p.nbytes + (map_p.nbytes if map_p is not None else 0)
    d2h_bytes = xout_p.nbytes

    s = S.SolverB200(device=local_rank, rank=rank, nranks=world, nccl_id=nccl_id)

    stage_t = {}

    def timed(name, fn, *a):
        t = time.perf_counter()
        r = fn(*a)
        stage_t[name] = stage_t.get(name, 0.0) + time.perf_counter() - t
        return r

    def upload_and_pattern():
        s.initialise(size_local, num.size_global)
        s.set_options(rtol=RTOL, max_it=100000, pc_type=S.PC_JACOBI)
        timed("set_mesh", s.set_mesh, kind, conn_p, coords_p, map_p)
        timed("set_pattern", s.set_pattern_nodal, nda_p)     # element dof lists are formed on the GPU
        timed("set_applied", s.set_applied, applied_p)

    def hot_step():
        s.setZero()
        s.assemble(ed, td)
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
    asm_b, spmv_b, cgit_b = algorithmic_bytes(conn.shape[1], m.nNode, size_local, nnz_local, m.dbc_node.size)
    spmv_gbs = spmv_b / spmv_avg / 1e9 if spmv_avg > 0 else 0.0
    asm_gbs = asm_b * args.steps / (t_asm if t_asm > 0 else 1) / 1e9
    cgit_gbs = cgit_b * its / t_solve / 1e9

    # FP64 pipe of the value pass (BASELINE.json north_star: "with the FP64 pipe reported for the element kernels"):
    # the row-gather kernel issues 114 FP64 instructions per (row, element) incidence (no FMA by design: bit-identical to
    # the reference's evaluation order), 4 incidences per tetrahedron; peak = SMs x 64 FP64 lanes x the SM clock under load.
    asm_fp64 = None
    asm_kernel = None
    try:
        asm_mode = s.assembly_mode()
        asm_kernel = {0: "assemble_kernel (row gather, binary-search slots)", 1: "assemble_sell_kernel (streamed row gather)",
                      2: f"assemble_tiled_kernel (compute-once tiles: {asm_mode[1]} tiles, {asm_mode[2]:.2f} visits/element)"}[asm_mode[0]]
        if asm_mode[0] != 1:
            raise RuntimeError("FP64 instruction count below is the row-gather kernel's")
        sm_count = torch.cuda.get_device_properties(local_rank).multi_processor_count
        mhz = float((clocks or {}).get("sm_mhz") or (clocks or {}).get("sm_max_mhz") or 1965.0)
        fp64_ops = 114.0 * 4.0 * conn.shape[1]
        fp64_peak = sm_count * 64 * mhz * 1e6
        fp64_rate = fp64_ops * args.steps / (t_asm if t_asm > 0 else 1)
        asm_fp64 = {"ops_per_launch": fp64_ops, "achieved_gops": fp64_rate / 1e9, "peak_gops": fp64_peak / 1e9,
                    "frac": fp64_rate / fp64_peak, "unit": "FP64 instr/s (DADD/DMUL, no FMA by design)",
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

    # ---- solution sanity on every run: nodally ~exact quadratic ----
    u = D.nodal_solution(num, xout_p)[0]
    err = float(np.abs(u - (m.coords ** 2).sum(0)).max())

    line = {
        "metric": "poisson_cg_dof_iter_per_s", "value": value, "unit": "DOF-iter/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"genTetra {args.cells}^3 x6 P1-tet Poisson on [-1,1]^3, Jacobi-CG rtol {RTOL:g}",
                   "elements": int(m.nElem), "nodes": int(m.nNode), "dof": int(N), "nnz": nnz_total,
                   "partition": "none" if world == 1 else args.partition, "parallelism": f"rows{world}",
                   "exchange": {0: "none", 1: "nccl", 2: "peer-memory kernels (NVLink)"}[s.comm_mode()],
                   "l2": "inputs larger than L2 (matrix >= 1.4 GB per pass vs 126 MB L2); no flush needed",
                   "timing": "CUDA events on the library stream (t_assemble, t_solve), max over ranks; ms_per_step = host wall between barriers"},
        "iterations_per_step": its_per_step, "reason": info["reason"], "max_nodal_error": err,
        "assembly": {"metric": "assembly_melem_per_s", "value": asm_value, "unit": "Melem/s",
                     "ms_per_pass": 1e3 * t_asm / args.steps,
                     "roofline": {"bound": "hbm (kernel is FP64-pipe/issue bound, see DESIGN.md)", "achieved": asm_gbs, "peak": peak,
                                  "unit": "GB/s", "frac": asm_gbs / peak,
                                  "traffic": (3.552e9 if (world == 1 and args.cells == 200) else None), "bytes_per_launch": asm_b, "peak_source": peak_src, "scope": "rank 0 share"},
                     "kernel": asm_kernel, "fp64_pipe": asm_fp64},
        "cg_iteration": {"ms_per_iteration": 1e3 * t_solve / max(its, 1), "achieved_gbs": cgit_gbs, "frac": cgit_gbs / peak,
                         "bytes_per_iteration": cgit_b},
        # dominant kernel = the persistent CG kernel (one cooperative launch per solve: set-up + every iteration);
        # duration = CUDA events on the launching stream around that launch, live in the timed steps above
        "roofline": {"bound": "hbm", "kernel": "cg_persistent_kernel", "achieved": cgit_gbs, "peak": peak, "unit": "GB/s",
                     "frac": cgit_gbs / peak,
                     # dram__bytes_read+write of this kernel from the committed ncu --set full capture (profiles/r01_ncu_final_c5.txt:
                     # 37.31 GB for set-up + 16 iterations of C5 on one GPU = 2.27 GB per iteration), scaled to this launch
                     "traffic": (2.27e9 * its_per_step if (world == 1 and args.cells == 200) else None),
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

    # ---- CPU baseline beside it (rank 0, N = 1 only): bounded sample of the same workload ----
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        from oracle import pyoracle as O
        threads = O.num_threads()
        rp, col, val = s.get_csr()
        rhs = s.get_rhs()
        ne = min(args.cpu_elems, m.nElem)
        mask = np.zeros(m.nElem, np.uint8)
        mask[:ne] = 1
        v2, r2 = np.zeros_like(val), np.zeros_like(rhs)
        t0 = time.perf_counter()
        O.assemble(O.POISSON_TETRA, num.conn_new, m.coords, None, num.elemDof, num.solnApplied, ed, td, rp, col, elem_mask=mask,
                   threads=threads, val=v2, rhs=r2)
        t1 = time.perf_counter()
        O.cg_jacobi(rp, col, val, rhs, rtol=RTOL, threads=threads, fixed_its=args.cpu_its)
        t2 = time.perf_counter()
        line["cpu_baseline"] = {
            "value": N * args.cpu_its / (t2 - t1), "unit": "DOF-iter/s", "cores": threads, "kind": "port",
            "assembly_melem_per_s": ne / (t1 - t0) / 1e6,
            "sample": f"first {ne} of {m.nElem} elements assembled ({t1 - t0:.1f} s); {args.cpu_its} CG iterations on the full system ({t2 - t1:.1f} s)",
            "note": "CPU restatement of the PETSc/MPI path (PETSc unavailable): OpenMP oracle on the box's host cores"}
    s.free()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
