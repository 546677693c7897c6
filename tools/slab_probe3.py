import sys
import numpy as np, torch
sys.path.insert(0, '.')
from stencils_b200 import _abi as A
from stencils_b200._desc import build_desc
from stencils_b200.stencils import Moore
dev = torch.device('cuda', 0)
W, H, N = 16384, 16384, 512
lib = A.lib(); offs = Moore(1).offsets(); st = torch.cuda.current_stream().cuda_stream
def timed(fn, n=N):
    fn(8); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(n); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
h = build_desc(size=(W, H), eltype=A.U8, out_eltype=A.U8, offsets=offs, radius=1, boundary=(A.WRAP, A.WRAP), reducer=A.LIFE, born_mask=8, survive_mask=12, flags=A.FLAG_CELLS_01)
bigA = torch.randint(0, 2, (H + 256, W), dtype=torch.uint8, device=dev); bigB = torch.zeros_like(bigA)
for sa, sb_ in ((0, 0), (1, 1), (1, 0), (0, 1), (16, 16), (128, 128), (128, 0), (64, 64)):
    x, y = bigA[sa:sa + H], bigB[sb_:sb_ + H]
    def loop(n):
        bufs = [x, y]
        for i in range(n):
            A.check(lib.sb200_gather(h.ptr(), bufs[i & 1].data_ptr(), bufs[1 - (i & 1)].data_ptr(), st))
    print(f'plain desc, src shifted {sa:3d} rows, dst shifted {sb_:3d} rows: us/step {timed(loop):.2f}  (A base mod 2MiB = {bigA.data_ptr() % (1<<21)}, B = {bigB.data_ptr() % (1<<21)})')
