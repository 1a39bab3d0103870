"""torchrun helper: per-solve times of the CG variants on N GPUs (diagnostics)"""
import os, sys, json, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dune_hdd_b200 as hdd
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
comm = hdd.parallel.init_comm(rank, world, lr)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
g = hdd.grids.cube(n, partitions=(8, 8))
roff = hdd.parallel.rank_cell_offsets(g, world)
d = hdd.BlockSWIPDG(g, hdd.problems.ESV2007(), device=lr, cell_range=(int(roff[rank]), int(roff[rank + 1])), comm=comm)
d.init()
out = {}
for typ in ("cg.mg", "cg.mg", "cg.mg", "cg.blockdiagonal"):
    maxit = 200000 if typ == "cg.mg" else 300
    try:
        t = time.time()
        u, info = d.uncached_solve({"type": typ, "precision": 1e-10, "max_iter": maxit}, return_info=True, copy_to_host=False)
        out.setdefault(typ, []).append((info["iterations"], info["seconds"], time.time() - t))
    except hdd.discretizations.linear_solver_failed as e:
        out.setdefault(typ, []).append(str(e)[:80])
if rank == 0:
    print(json.dumps(out))
dist.barrier()
dist.destroy_process_group()
