"""torchrun check of the sharded rollout on real GPUs: each rank plays its contiguous env range,
one NCCL all-gather combines the 24-word statistics, and the result must equal the single-GPU
run of the whole range (Philox is keyed by the global env id)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, '.')
from simulate_2048_b200 import ops  # noqa: E402
from simulate_2048_b200.parallel import EpisodeStats, shard_range, sharded_rollout  # noqa: E402

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
n, seed = 1 << 21, 2048
out, stats = sharded_rollout(n, seed=seed, per_env=True, device=torch.device('cuda', local))
whole = ops.rollout(n, seed=seed, per_env=True, device=torch.device('cuda', local))
ref = EpisodeStats.from_vector(whole.stats.cpu())
lo, hi = shard_range(n, rank, world)
ok = (stats == ref) and torch.equal(out.score, whole.score[lo:hi]) and torch.equal(out.boards, whole.boards[lo:hi])
flag = torch.tensor([1 if ok else 0], device='cuda')
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print('sharded == single-GPU:', bool(flag.item()), stats.summary(), file=sys.stderr)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
