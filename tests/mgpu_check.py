"""torchrun helper: N-rank fused DSM == single-rank fused DSM, bit for bit, through both exchange transports (the
checks live in vissatsatellitestereo_b200/mgpu_selfcheck.py; bench.py --gpus N runs them too).  By hand:
python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tests/mgpu_check.py"""
import os
import sys

import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    from vissatsatellitestereo_b200 import mgpu_selfcheck
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    local_rank = int(os.environ.get('LOCAL_RANK', rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    dist.init_process_group('nccl', device_id=dev)
    res = mgpu_selfcheck.run(dev, rank, world, log=lambda *a: print('[rank {}]'.format(rank), *a, flush=True))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if res['bit_identical'] else 1)


if __name__ == '__main__':
    main()
