#!/usr/bin/env python
"""Multi-rank version of tools/pcg_trace.py (needs a -DPFEM_PCG_TRACE build): stage timing of the persistent CG kernel's last
iteration on every rank (CTA 0, thread 0, clock64).
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/pcg_trace_mp.py [cells] [metis|slab]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from pfemfort_b200 import driver as D, mesh as M, solver as S  # noqa: E402
from tools.pcg_trace import NAMES  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group(backend="gloo", rank=rank, world_size=world)
cells = int(sys.argv[1]) if len(sys.argv) > 1 else 100
part = sys.argv[2] if len(sys.argv) > 2 else "metis"
m = M.gen_tetra(-1, 1, cells, -1, 1, cells, -1, 1, cells)
npart = torch.zeros(m.nNode, dtype=torch.int32)
if rank == 0:
    p = D.partition(m, S.POISSON_TETRA, world)[1] if part == "metis" else (np.arange(m.nNode, dtype=np.int64) * world // m.nNode).astype(np.int32)
    npart = torch.from_numpy(np.ascontiguousarray(p, dtype=np.int32))
dist.broadcast(npart, 0)
num = D.number(m, S.POISSON_TETRA, world, npart.numpy())
idt = torch.zeros(128, dtype=torch.uint8)
if rank == 0:
    idt = torch.frombuffer(bytearray(S.comm_unique_id()), dtype=torch.uint8).clone()
dist.broadcast(idt, 0)
s = S.SolverB200(device=int(os.environ.get("LOCAL_RANK", rank)), rank=rank, nranks=world, nccl_id=bytes(idt.numpy().tobytes()))
for _ in range(2):
    info = D.run_rank(s, m, num, rank=rank, rtol=1e-30, max_it=300)
out = np.zeros(64, np.int64)
rc = S.load_library().pfem_debug_pcg_trace(out.ctypes.data_as(C.POINTER(C.c_longlong)))
assert rc == 0, "not a PFEM_PCG_TRACE build"
lines = [f"rank {rank}: rows {num.row_range(rank)[1] - num.row_range(rank)[0]}  {1e6 * info['t_solve'] / info['its']:.2f} us/iteration over {info['its']} iterations ({part})"]
keys = sorted(k for k in NAMES if out[k] > 0)
prev = t0 = out[0]
for k in keys:
    lines.append(f"  [{k:2d}] {NAMES[k]:42s} +{out[k] - prev:7d} cyc   (t = {out[k] - t0:7d})")
    prev = out[k]
allr = [None] * world
dist.all_gather_object(allr, "\n".join(lines))
if rank == 0:
    print("\n".join(allr), flush=True)
s.free()
dist.barrier()
dist.destroy_process_group()
