"""Where does the slab iterator's per-step overhead come from? 1 GPU, Life 16384^2 (run under gpurun)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
import stencils_b200 as sb
from stencils_b200 import _abi as A
from stencils_b200.slab import SlabIterator
from stencils_b200.synth import synth_torch
from stencils_b200.stencils import Moore

dev = torch.device('cuda', 0)
shape = (16384, 16384)
field = synth_torch(shape, np.uint8, 0x5EED0002, dev)
S = sb.SwitchingStencilArray(field, sb.Moore(1), boundary=sb.Wrap())
def t_iter(n):
    global S
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); S = sb.iterate_(sb.Life(), S, n); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
t_iter(20)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
print("sb200_iterate          ms/step", t_iter(N))
t = field.permute(1, 0).contiguous()
for G in ((1, 16, 64) if N > 100 else (16,)):
    it = SlabIterator(t, offsets=Moore(1).offsets(), radius=1, reducer=A.LIFE, boundary=(A.WRAP, A.WRAP), eltype=A.U8, ghost=G,
                      rank=0, world=1, reducer_kwargs=dict(born_mask=8, survive_mask=12))
    it.step(2 * G + 4); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter(); e0.record(); it.step(N); e1.record(); w1 = time.perf_counter(); torch.cuda.synchronize()
    print(f'slab world=1 G={G:3d}     ms/step', e0.elapsed_time(e1) / N, ' host enqueue ms/step', (w1 - w0) / N * 1e3, A.lib().sb200_last_kernel().decode())
    del it
