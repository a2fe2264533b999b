"""debug: bench.multi_gpu_parity with two gloo ranks sharing cuda:0 (run under torchrun --nproc-per-node 2)"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import bench

rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(0)
dev = torch.device('cuda', 0)
dist.init_process_group('gloo')
import config
config.init()
arm = bench.Arm(dev, rank, world, shard=(sys.argv[2] if len(sys.argv) > 2 else 'batch'))
warm = int(sys.argv[1]) if len(sys.argv) > 1 else 0
if warm:
    arm.measure('train', 64, 2, 1)
    arm.release()
res = bench.multi_gpu_parity(arm)
if rank == 0:
    print({k: v for k, v in res.items() if k in ('worst_rel', 'dloss')})
dist.destroy_process_group()
