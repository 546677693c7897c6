"""Isolate the slab overhead: same kernel driven (a) by sb200_iterate, (b) by a Python loop of sb200_gather on the plain
descriptor, (c) by a Python loop on the ghost-plane descriptor with a fixed region, (d) with the shrinking regions."""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from stencils_b200 import _abi as A
from stencils_b200._desc import build_desc
from stencils_b200.synth import synth_torch
from stencils_b200.stencils import Moore
dev = torch.device('cuda', 0)
W, H, G = 16384, 16384, 16
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
lib = A.lib()
offs = Moore(1).offsets()
st = torch.cuda.current_stream().cuda_stream
def timed(fn, n=N):
    fn(8); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(n); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
a = synth_torch((W, H), np.uint8, 1, dev).permute(1, 0).contiguous(); b = torch.empty_like(a)
h = build_desc(size=(W, H), eltype=A.U8, out_eltype=A.U8, offsets=offs, radius=1, boundary=(A.WRAP, A.WRAP), reducer=A.LIFE, born_mask=8, survive_mask=12, flags=A.FLAG_CELLS_01)
print('a iterate (C loop)              us/step', timed(lambda n: A.check(lib.sb200_iterate(h.ptr(), a.data_ptr(), b.data_ptr(), n, st))))
def loop_plain(n):
    bufs = [a, b]
    for i in range(n):
        A.check(lib.sb200_gather(h.ptr(), bufs[i & 1].data_ptr(), bufs[1 - (i & 1)].data_ptr(), st))
print('b python loop, plain desc       us/step', timed(loop_plain))
ext = H + 2 * G
ga = torch.zeros((ext, W), dtype=torch.uint8, device=dev); gb = torch.zeros_like(ga)
def mk(lo, hi):
    return build_desc(size=(W, ext - 2), eltype=A.U8, out_eltype=A.U8, offsets=offs, radius=1, boundary=(A.WRAP, A.USE), reducer=A.LIFE,
                      src_off=(0, 1), dst_off=(0, 1), src_ext=(W, ext), dst_ext=(W, ext), region=((0, lo - 1, 0), (W, hi - 1, 0)),
                      flags=A.FLAG_CELLS_01, born_mask=8, survive_mask=12)
hg = mk(1, ext - 1)
def loop_ghost(n):
    bufs = [ga, gb]
    for i in range(n):
        A.check(lib.sb200_gather(hg.ptr(), bufs[i & 1].data_ptr(), bufs[1 - (i & 1)].data_ptr(), st))
print('c python loop, ghost desc fixed us/step', timed(loop_ghost))
hs = [mk(s, ext - s) for s in range(1, G + 1)]
def loop_shrink(n):
    bufs = [ga, gb]
    for i in range(n):
        A.check(lib.sb200_gather(hs[i % G].ptr(), bufs[i & 1].data_ptr(), bufs[1 - (i & 1)].data_ptr(), st))
print('d python loop, shrinking region us/step', timed(loop_shrink))
hh = mk(G, ext - G)
def loop_owned(n):
    bufs = [ga, gb]
    for i in range(n):
        A.check(lib.sb200_gather(hh.ptr(), bufs[i & 1].data_ptr(), bufs[1 - (i & 1)].data_ptr(), st))
print('e python loop, owned rows only  us/step', timed(loop_owned))
# (f) plain WRAP descriptor, buffers 2^28 + 512 KiB apart (the ghost layout's distance)
big2 = torch.zeros((2 * 16384 + 32, W), dtype=torch.uint8, device=dev)
fa, fb = big2[:16384], big2[16416:16416 + 16384]
def loop_f(n):
    bufs = [fa, fb]
    for i in range(n):
        A.check(lib.sb200_gather(h.ptr(), bufs[i & 1].data_ptr(), bufs[1 - (i & 1)].data_ptr(), st))
print('f plain desc, 2^28+512K apart   us/step', timed(loop_f), 'distance', fb.data_ptr() - fa.data_ptr())
# (g) ghost descriptor, parents exactly 2^28 bytes apart (ext = 16384 rows)
big = torch.zeros((2 * 16384, W), dtype=torch.uint8, device=dev)
g2a, g2b = big[:16384], big[16384:]
ext2 = 16384
hg2 = build_desc(size=(W, ext2 - 2), eltype=A.U8, out_eltype=A.U8, offsets=offs, radius=1, boundary=(A.WRAP, A.USE), reducer=A.LIFE,
                 src_off=(0, 1), dst_off=(0, 1), src_ext=(W, ext2), dst_ext=(W, ext2), flags=A.FLAG_CELLS_01, born_mask=8, survive_mask=12)
def loop_g(n):
    bufs = [g2a, g2b]
    for i in range(n):
        A.check(lib.sb200_gather(hg2.ptr(), bufs[i & 1].data_ptr(), bufs[1 - (i & 1)].data_ptr(), st))
print('g ghost desc, 2^28 apart        us/step', timed(loop_g), 'distance', g2b.data_ptr() - g2a.data_ptr(), 'plain distance', b.data_ptr() - a.data_ptr())
# (h) ring on axis 1 read straight through (off = 1) but boundary WRAP instead of USE
hh2 = build_desc(size=(W, ext - 2), eltype=A.U8, out_eltype=A.U8, offsets=offs, radius=1, boundary=(A.WRAP, A.WRAP), reducer=A.LIFE,
                 src_off=(0, 1), dst_off=(0, 1), src_ext=(W, ext), dst_ext=(W, ext), flags=A.FLAG_CELLS_01, born_mask=8, survive_mask=12)
def loop_h(n):
    bufs = [ga, gb]
    for i in range(n):
        A.check(lib.sb200_gather(hh2.ptr(), bufs[i & 1].data_ptr(), bufs[1 - (i & 1)].data_ptr(), st))
lib.sb200_launch_count(1)
print('h ring desc, WRAP not USE       us/step', timed(loop_h), 'launches', lib.sb200_launch_count(1))
# (i) ghost desc over random data instead of zeros
ga.copy_(torch.randint(0, 2, ga.shape, dtype=torch.uint8, device=dev))
print('i ghost desc, random cells      us/step', timed(loop_ghost))
# (j) plain desc over zeros
a.zero_(); b.zero_()
print('j plain desc, zero cells        us/step', timed(loop_plain))
a.copy_(torch.randint(0, 2, a.shape, dtype=torch.uint8, device=dev))
for lo, hi in ((1, H), (0, H - 1), (16, H), (0, H - 16), (7, H - 9)):
    hk = build_desc(size=(W, H), eltype=A.U8, out_eltype=A.U8, offsets=offs, radius=1, boundary=(A.WRAP, A.WRAP), reducer=A.LIFE,
                    born_mask=8, survive_mask=12, flags=A.FLAG_CELLS_01, region=((0, lo, 0), (W, hi, 0)))
    def loop_k(n):
        bufs = [a, b]
        for i in range(n):
            A.check(lib.sb200_gather(hk.ptr(), bufs[i & 1].data_ptr(), bufs[1 - (i & 1)].data_ptr(), st))
    print(f'k plain desc, region rows [{lo},{hi})  us/step', timed(loop_k))
for off in (2, 8, 16):
    e3 = H + 2 * off
    ra = torch.zeros((e3, W), dtype=torch.uint8, device=dev); rb = torch.zeros_like(ra)
    hr = build_desc(size=(W, H), eltype=A.U8, out_eltype=A.U8, offsets=offs, radius=1, boundary=(A.WRAP, A.USE), reducer=A.LIFE,
                    src_off=(0, off), dst_off=(0, off), src_ext=(W, e3), dst_ext=(W, e3), flags=A.FLAG_CELLS_01, born_mask=8, survive_mask=12)
    def loop_r(n):
        bufs = [ra, rb]
        for i in range(n):
            A.check(lib.sb200_gather(hr.ptr(), bufs[i & 1].data_ptr(), bufs[1 - (i & 1)].data_ptr(), st))
    print(f'l ring desc off={off:2d}, H rows        us/step', timed(loop_r))
    del ra, rb
