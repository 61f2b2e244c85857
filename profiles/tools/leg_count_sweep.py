import sys, os, json, numpy as np, torch
sys.path.insert(0, os.getcwd())
import igate4xsoftphonedsp_b200 as ig
from igate4xsoftphonedsp_b200 import synth
vp = ig.VoicePath(0); vp.use_torch_stream()
dev = torch.device("cuda", 0)
for G in (4, 2, 1, 3, 8, 32):
    Cn = 4096; B = Cn // G if Cn % G == 0 else 1365; Cn = B * G; F = 1640
    g = torch.Generator(device=dev).manual_seed(1)
    codes = torch.randint(0, 256, (F, Cn, 160), dtype=torch.uint8, device=dev, generator=g)
    law = torch.from_numpy(synth.laws(Cn)).to(dev)
    gain = torch.zeros((F, B, G), dtype=torch.int16, device=dev); gain[:, :, :(int(os.environ['IGD_SWEEP_OPEN']) if os.environ.get('IGD_SWEEP_OPEN') else max(1, G // 2))] = 256   # default: half of the legs open
    gain = gain.reshape(F, Cn).contiguous()
    out_law = torch.from_numpy(synth.out_laws(B)).to(dev)
    out = vp.alloc_outputs(F, B, G)
    f = lambda: vp.process_batch(codes, law, gain, out_law, G, out=out)
    f(); f(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
    nbytes = F * B * (G * 160 + 480 + G * 16)          # SURVEY 8(d) accounting (bridge record not counted)
    print(f"G={G} B={B}: {best:.3f} ms  {nbytes / best / 1e6:.0f} GB/s  {nbytes / best / 1e6 / 6552.6:.1%}")
