"""Where does the data-parallel exchange run? CUDA events on the main and the communication stream around one critic
sub-step's exchange (torchrun, N ranks): times in microseconds relative to the end of the sub-step's last kernel."""
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
B = 128
hp = make_hparams(B)
g, d = get_models(hp, None)
gan = get_algorithm(hp, g, d, None)
eng = gan.engine
real = torch.from_numpy(np.random.RandomState(rank).uniform(0, 1, (B, 2048, 102)).astype(np.float32)).cuda()
for _ in range(3):
  gan.train(real)
gan._peer_setup(dist)
main = torch.cuda.current_stream()
ev = lambda: torch.cuda.Event(enable_timing=True)
rows = []
for it in range(4):
  e = {k: ev() for k in ('start', 'step', 'prefetch', 'waited', 'adam', 'c_wait', 'c_b0', 'c_red', 'c_b1', 'c_ag')}
  torch.cuda.synchronize(); dist.barrier()
  e['start'].record()
  eng.critic_step(real, update=False, sync=False, same_real=it > 0, want_fake32=False)
  e['step'].record()
  comm = gan._comm_stream if hasattr(gan, '_comm_stream') else torch.cuda.Stream()
  gan._comm_stream = comm
  mode = os.environ.get('CG_DP_COMM', 'p2p')
  with torch.cuda.stream(comm):
    eng.stream_wait_bucket(L.DISCRIMINATOR, 2, comm)
    e['c_wait'].record()
    if mode == 'p2p':
      buf, hdl, ptrs, red, rptrs = gan._peer[L.DISCRIMINATOR]
      hdl.barrier(channel=0); e['c_b0'].record()
      if red is None:
        eng.reduce_peer_grads(L.DISCRIMINATOR, ptrs, comm); e['c_red'].record()
        hdl.barrier(channel=1); e['c_b1'].record(); e['c_ag'].record()
      else:
        eng.peer_reduce_scatter(L.DISCRIMINATOR, ptrs, rank, comm); e['c_red'].record()
        hdl.barrier(channel=1); e['c_b1'].record()
        eng.peer_all_gather(L.DISCRIMINATOR, rptrs, rank, comm); e['c_ag'].record()
    else:
      e['c_b0'].record()
      w = dist.all_reduce(eng.grad_tensor(L.DISCRIMINATOR), async_op=True)
      w.wait(); e['c_red'].record(); e['c_b1'].record(); e['c_ag'].record()
  t_host0 = __import__('time').time()
  eng.prefetch_generator(real, for_generator_step=False, want_fake32=False)
  e['prefetch'].record()
  main.wait_stream(comm)
  e['waited'].record()
  (eng.apply_update_reduced if mode == 'p2p' else eng.apply_update)(L.DISCRIMINATOR)
  e['adam'].record()
  torch.cuda.synchronize()
  eng.critic_step(real, update=False, sync=True, same_real=True, want_fake32=False, gen_prefetched=True)   # consume the prefetch
  t = {k: e['step'].elapsed_time(v) * 1e3 for k, v in e.items()}
  rows.append(t)
if rank == 0:
  for t in rows[1:]:
    print('%s: step at 0 (took %.0f us) | main: prefetch done %+.0f, comm waited %+.0f, adam done %+.0f | comm: event %+.0f, barrier0 %+.0f, reduce %+.0f, barrier1 %+.0f, gather %+.0f'
          % (os.environ.get('CG_DP_COMM', 'p2p'), -t['start'], t['prefetch'], t['waited'], t['adam'], t['c_wait'], t['c_b0'], t['c_red'], t['c_b1'], t['c_ag']))
dist.barrier()
dist.destroy_process_group()
