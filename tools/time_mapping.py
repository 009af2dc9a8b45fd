"""Time G.mapping (label-map Encoder + z/c mapping network) next to G.synthesis at config 2."""
import sys, time
import torch
sys.path.insert(0, '.')
from pix2pix3d_b200 import configs

name = 'seg2cat_512'
w = configs.WORKLOADS[name]
B = w['batch']
G = configs.build_generator(name, seed=0, device='cuda', with_mapping=True)
z = torch.randn(B, 512, device='cuda')
c = configs.camera_labels(B, seed=1, preset=w['preset']).cuda()
mask = configs.label_map(name, B, seed=2).cuda()
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


with torch.no_grad():
    t_map = timeit(lambda: G.mapping(z, c, {'mask': mask, 'pose': c}))
    ws = G.mapping(z, c, {'mask': mask, 'pose': c})
    t_syn = timeit(lambda: G.synthesis(ws, c, noise_mode='const', neural_rendering_resolution=w['nrr']))
print(f'mapping {t_map:.2f} ms   synthesis (eager) {t_syn:.2f} ms   batch {B}')
from torch.profiler import profile, ProfilerActivity
with torch.no_grad(), profile(activities=[ProfilerActivity.CUDA]) as prof:
    G.mapping(z, c, {'mask': mask, 'pose': c})
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=12, max_name_column_width=60))
