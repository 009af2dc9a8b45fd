"""Launch the fused render kernel alone at BASELINE config-2 size (for ncu captures and quick timing)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pix2pix3d_b200 import native, configs
from pix2pix3d_b200.training.triplane_cond import OSGDecoder_semantic_lateSeparate

def main(reps=3, impl='auto', B=4, H=256, nrr=128, Sc=48, Sf=48):
    dev = torch.device('cuda')
    torch.manual_seed(0)
    planes = torch.randn(B, 3, 32, H, H, device=dev)
    dec_m = OSGDecoder_semantic_lateSeparate(32, {'decoder_lr_mul': 1, 'decoder_output_dim': 32, 'sigmoid': False,
                                                  'semantic_channels': 6}).to(dev).requires_grad_(False)
    c = configs.camera_labels(B, 0).to(dev)
    o, d = native.ray_sampler(c[:, :16].reshape(-1, 4, 4), c[:, 16:].reshape(-1, 3, 3), nrr)
    R = nrr * nrr
    dc = torch.linspace(2.25, 3.3, Sc, device=dev).reshape(1, 1, Sc) + torch.rand(B, R, Sc, device=dev) * ((3.3 - 2.25) / (Sc - 1))
    u = torch.rand(B * R, Sf, device=dev)
    dec = native.pack_decoder(dec_m)
    pcl = native.planes_to_channels_last(planes)
    for _ in range(2):
        native.render_fwd(pcl, dec, o, d, dc, u, 1.0, impl=impl)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        native.render_fwd(pcl, dec, o, d, dc, u, 1.0, impl=impl)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    gb = configs.render_algorithmic_bytes(B, R, Sc + Sf) / 1e9
    print(f'render_fwd[{impl}]: {ms:.3f} ms  {gb / ms * 1e3:.1f} GB/s touched  ({B * R / ms * 1e3:.3e} rays/s)')

if __name__ == '__main__':
    # usage: profile_render.py [reps] [impl] [B nrr Sc Sf]   e.g. `10 tc_pairs 8 64 64 64` for the config-4 sampling
    extra = [int(v) for v in sys.argv[3:7]]
    kw = dict(zip(('B', 'nrr', 'Sc', 'Sf'), extra))
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 3, sys.argv[2] if len(sys.argv) > 2 else 'auto', **kw)
