"""Probe for tests/test_cpusim.py: makes the simulator's detectors fire on purpose.  TEST INFRASTRUCTURE."""
import os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE))); sys.path.insert(0, HERE)
import simtorch; simtorch.install()
import torch, torch.distributed as dist, candmc_b200 as cb
from candmc_b200._lib import lib, check
rank = int(os.environ["RANK"])
dist.init_process_group("nccl")
world = cb.init_world(rank, 2, rank)
x = torch.zeros(4, dtype=torch.float64, device="cuda")
if sys.argv[1] == "mismatch":     # rank 0 broadcasts while rank 1 all-reduces on the same communicator
    if rank == 0:
        check(lib().candmc_comm_bcast(world.cm, x.data_ptr(), 4, 0, None))
    check(lib().candmc_comm_allreduce_sum(world.cm, x.data_ptr(), x.data_ptr(), 4, None))
else:                             # rank 1 never joins: the all-reduce of rank 0 can not complete
    if rank == 0:
        check(lib().candmc_comm_allreduce_sum(world.cm, x.data_ptr(), x.data_ptr(), 4, None))
    else:
        import time; time.sleep(5)
