"""e2e (host buffers, H2D + sweep + D2H per step) against the raw PCIe copy rates of the box."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, '.')
import stencils_b200 as sb
from stencils_b200.synth import synth_np
shape = (16384, 16384)
host = torch.empty(tuple(reversed(shape)), dtype=torch.uint8, pin_memory=True)
host.copy_(torch.from_numpy(np.ascontiguousarray(synth_np(shape, np.uint8, 0x5EED0002).T)))
out = torch.empty_like(host, pin_memory=True)
d = torch.empty_like(host, device='cuda'); d2 = torch.empty_like(d)
def t(fn, n=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print('H2D 256 MiB ms', t(lambda: d.copy_(host, non_blocking=True)))
print('D2H 256 MiB ms', t(lambda: out.copy_(d, non_blocking=True)))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1): d.copy_(host, non_blocking=True)
    with torch.cuda.stream(s2): out.copy_(d2, non_blocking=True)
print('H2D || D2H ms', t(both))
a = sb.StencilArray(host.numpy().T, sb.Moore(1), boundary=sb.Wrap()); dst = out.numpy().T
print('e2e ms/step', t(lambda: sb.mapstencil_(sb.Life(), dst, a)), 'chunks', os.environ.get('SB200_HOST_CHUNKS', '16'))
