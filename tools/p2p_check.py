"""Peer-memory gradient reduction vs NCCL all-reduce on the same gradients (run under torchrun, 2/4/8 ranks)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from bench import make_hparams
from calciumgan_b200 import _lib as L
from calciumgan_b200.algorithms.registry import get_algorithm
from calciumgan_b200.models.registry import get_models

rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
dist.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ['LOCAL_RANK'])))
B = 8
hp = make_hparams(B)
g, d = get_models(hp, None)
gan = get_algorithm(hp, g, d, None)
eng = gan.engine
assert gan._peer_setup(dist), 'peer setup failed'
real = torch.from_numpy(np.random.RandomState(rank).uniform(0, 1, (B, 2048, 102)).astype(np.float32)).cuda()
for which, step in ((L.DISCRIMINATOR, lambda: eng.critic_step(real, update=False)), (L.GENERATOR, lambda: eng.generator_step(real, update=False))):
  for it in range(3):
    step()
    buf, hdl, ptrs, red_buf, rptrs = gan._peer[which]
    n = eng.num_params(which)
    own = torch.as_tensor(np.concatenate([x.ravel() for x in eng.get_grads(which)])).cuda()
    print('rank %d which %d it %d: |buf - get_grads| %.3e, buf norm %.4e, ptr ok %s' % (
        rank, which, it, float((buf[:n] - own).abs().max()), float(buf[:n].norm()), ptrs[rank] == buf.data_ptr()))
    ref = own.clone()
    dist.all_reduce(ref)
    works = gan._allreduce_start(dist, which)
    assert works == 'peer'
    torch.cuda.current_stream().wait_stream(gan._comm_stream)
    torch.cuda.synchronize()
    red = eng.reduced_grad_tensor(which)
    err = float((red - ref).abs().max() / ref.abs().max())
    peer_view = hdl.get_buffer((rank + 1) % world, (n,), torch.float32)
    print('rank %d which %d it %d: reduced vs nccl rel err %.3e; peer norm via get_buffer %.4e' % (rank, which, it, err, float(peer_view.norm())))
dist.barrier()
dist.destroy_process_group()
