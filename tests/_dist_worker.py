"""Worker for test_gather_scores_gloo_world2 (launched by torch.distributed.run, gloo backend, CPU)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vtamiq_b200.parallel import gather_scores, shard_pairs  # noqa: E402

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n = 7
a, b = shard_pairs(n, rank, world)
local = torch.arange(a, b, dtype=torch.float32) * 1.5 + 0.25      # stands in for this rank's scores
full = gather_scores(local, n)
assert torch.equal(full, torch.arange(n, dtype=torch.float32) * 1.5 + 0.25), full
dist.barrier()
print("DIST_OK", rank)
dist.destroy_process_group()
