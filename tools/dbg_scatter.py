import numpy as np, sys
sys.path.insert(0, '.')
from stencils_b200 import _abi as A
from stencils_b200._desc import build_desc
from oracle import np_restatement as npr
from tests.util import gpu_scatter
rng = np.random.default_rng(1)
for offs, R, (ny, nx) in [([(-1, 1), (-2, -1), (1, 0), (-2, 2)], 2, (2600, 75)), (npr.offsets("Window", 1, 2), 1, (1028, 300)), (npr.offsets("Circle", 3, 2), 3, (1100, 63))]:
    for dt in (np.float32, np.float64):
        src = np.asfortranarray(rng.random((ny, nx)).astype(dt)); w = rng.random(len(offs)).astype(dt)
        et = A.ELTYPE_OF_DTYPE[np.dtype(dt)]
        for op, rule, flags in [(A.OP_ADD, A.SCATTER_CENTER_WEIGHTS, 0), (A.OP_MAX, A.SCATTER_CENTER_WEIGHTS, A.FLAG_ZERO_DEST), (A.OP_MIN, A.SCATTER_WEIGHTS, 0)]:
            h = build_desc(size=(ny, nx), eltype=et, out_eltype=et, offsets=offs, radius=R, boundary=A.REMOVE, weights=w, scatter_op=op, scatter_rule=rule, flags=flags)
            gpu_scatter(h, src, np.zeros_like(src, order="F"))
            print(len(offs), R, ny, nx, dt.__name__, op, rule, flags, A.lib().sb200_last_kernel().decode(), A.lib().sb200_last_error().decode())
