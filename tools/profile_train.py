"""Top CUDA kernels of the config-5 training iteration (product arm) by total device time: `python tools/profile_train.py [ours|reference]`."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'baseline'))
import torch
import ref_harness as rh
arm = sys.argv[1] if len(sys.argv) > 1 else 'ours'
if arm == 'ours':
    import pix2pix3d_b200
    pix2pix3d_b200.install(reference_root=rh.REF_ROOT)
else:
    rh.import_reference()
from pix2pix3d_b200 import train_step as ts
dev = torch.device('cuda')
cfg = dict(ts.AFHQ_TRAIN)
st = ts.build(cfg, dev)
batch = ts.synthetic_batch(cfg, dev, 100)
for _ in range(3):
    ts.run_iteration(st, batch)
st.batch_idx = 1                     # a main-phases-only iteration
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    ts.run_iteration(st, batch)
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages():
    t = getattr(e, 'device_time_total', None) or getattr(e, 'cuda_time_total', 0)
    if t > 0 and e.device_type.name == 'CUDA':
        rows.append((t, e.count, e.key))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print(f'arm {arm}: device time of one main-phase iteration {tot / 1000:.1f} ms over {sum(r[1] for r in rows)} kernel launches')
for t, n, k in rows[:40]:
    print(f'{t / 1000:9.2f} ms {100 * t / tot:5.1f}% {n:5d}  {k[:150]}')
