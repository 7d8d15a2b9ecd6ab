"""torchrun --nproc-per-node 2: token-sharded factorisation (Gram all-reduce over NCCL) against the
single-GPU factorisation of the whole matrix. Run on a 2-GPU box: gpurun --gpus 2."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from xkv_b200 import factorize, parallel, synthetic

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
S, n, r = 8192, 2048, 256
x = synthetic.group_matrix(S, n, 1.0, seed=42, device=dev)      # same seed: identical on every rank
b, e = parallel.token_shard(S, world, rank)
(f_shard,) = factorize.factorize_batch([x[b:e].contiguous()], r, process_group=dist.group.WORLD)
(f_full,) = factorize.factorize_batch([x], r)
torch.cuda.synchronize()
xd = x.double()
err_full = (torch.linalg.norm(xd - f_full.reconstruct().double()) / torch.linalg.norm(xd)).item()
rec_local = f_shard.reconstruct().double()
num = torch.linalg.norm(xd[b:e] - rec_local) ** 2
dist.all_reduce(num)
err_shard = (num.sqrt() / torch.linalg.norm(xd)).item()
same_v = (f_shard.Vt.float() - f_full.Vt.float()).abs().max().item()
ok = err_shard <= 1.002 * err_full
if rank == 0:
    print(f"token-sharded x{world}: rel err {err_shard:.6f} vs single-GPU {err_full:.6f}; max|Vt diff| {same_v:.3e}; "
          f"{'OK' if ok else 'FAIL'}")
dist.destroy_process_group()
sys.exit(0 if ok else 1)
