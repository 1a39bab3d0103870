"""torchrun helper: HDD_TIMING=1 phases of mesh / discretization set-up on N GPUs (diagnostics)"""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dune_hdd_b200 as hdd
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
comm = hdd.parallel.init_comm(rank, world, lr)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
g = hdd.grids.cube(n, partitions=(8, 8), pinned=True)
roff = hdd.parallel.rank_cell_offsets(g, world)
for k in range(3):
    dist.barrier()
    t = time.time()
    d = hdd.BlockSWIPDG(g, hdd.problems.ESV2007(), device=lr, cell_range=(int(roff[rank]), int(roff[rank + 1])), comm=comm)
    t1 = time.time()
    d.init()
    t2 = time.time()
    sys.stderr.write("[rank %d pass %d] create %.3f s, init %.3f s, cpus %d\n" % (rank, k, t1 - t, t2 - t1, os.cpu_count()))
    del d
dist.barrier()
dist.destroy_process_group()
