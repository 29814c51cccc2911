"""GPU debug helper: one stem_tc step; prints which 128-row tiles / columns differ from the reference."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
g.build()
import tenet_jl_b200 as tb

def crand(rng, shape):
    return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(np.complex64)

ctx = tb.default_context(0)
for (N, K) in [(16, 16), (64, 16), (32, 128)]:
    rng = np.random.default_rng(N * 131 + K)
    a = crand(rng, (2,) * 17 + (K,))
    b = crand(rng, (N, K))
    big = [f"m{i}" for i in range(17)]
    c = tb.binary_einsum(tb.Tensor(a, big + ["k"]), tb.Tensor(b, ["n", "k"]))
    ref = np.tensordot(a.astype(np.complex128), b.astype(np.complex128), axes=([17], [1]))
    got = np.reshape(c.parent, (-1, N), order="F")
    r = np.reshape(ref, (-1, N), order="F")
    bad = np.abs(got - r) > 1e-3 * np.abs(r).max()
    tiles = np.unique(np.nonzero(bad)[0] // 128)
    print(f"N={N} K={K} kernel={ctx.last_kernel} bad {bad.sum()} of {bad.size}; bad tiles {len(tiles)} of {got.shape[0] // 128}; "
          f"tile parity in CTA order (tile//148 % 2): {np.unique((tiles // 148) % 2)}; first tiles {tiles[:12]}; bad cols {np.unique(np.nonzero(bad)[1])[:20]}")
    if bad.any():
        t = tiles[0]
        blk_g, blk_r = got[t * 128:(t + 1) * 128], r[t * 128:(t + 1) * 128]
        # is the bad tile equal to some OTHER tile of the reference?
        for cand in (t - 148, t + 148, t - 1, t + 1):
            if 0 <= cand < got.shape[0] // 128:
                d = np.abs(blk_g - r[cand * 128:(cand + 1) * 128]).max() / np.abs(r).max()
                print(f"   tile {t} vs reference tile {cand}: {d:.2e}")
        print("   rows bad in tile", np.unique(np.nonzero(bad[t * 128:(t + 1) * 128])[0])[:16], " zeros?", float(np.abs(blk_g).max()))
