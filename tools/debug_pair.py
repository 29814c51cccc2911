"""GPU debug helper: run one c64 GEMM-shaped einsum with the CTA-pair kernel and the 1-CTA kernel, locate mismatches."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
g.build()
import tenet_jl_b200 as tb

def crand(rng, shape):
    return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(np.complex64)

ctx = tb.default_context(0)
for (lm, N, K) in [(17, 256, 512), (17, 128, 256), (15, 256, 512), (17, 256, 128), (13, 256, 512)]:
    rng = np.random.default_rng(N + K)
    a = crand(rng, (2,) * lm + (K,))
    b = crand(rng, (N, K))
    big = [f"m{i}" for i in range(lm)]
    ta, tb_ = tb.Tensor(a, big + ["k"]), tb.Tensor(b, ["n", "k"])
    ref = np.tensordot(a.astype(np.complex128), b.astype(np.complex128), axes=([lm], [1])).reshape(-1, N)
    for out in [None, ["n"] + big]:
        res = {}
        for pair in (0, 1):
            ctx.set_option(tb._lib.TNB_OPT_GEMM_PAIR, pair)
            c = tb.binary_einsum(ta, tb_, out=out)
            r = c.parent
            r = r.reshape(-1, N) if out is None else r.reshape(N, -1).T
            res[pair] = r.copy()
        ctx.set_option(tb._lib.TNB_OPT_GEMM_PAIR, 1)
        mx = np.abs(ref).max()
        e0, e1 = np.abs(res[0] - ref) / mx, np.abs(res[1] - ref) / mx
        bad = e1 > 1e-4
        print(f"M=2^{lm} N={N} K={K} out={'default' if out is None else 'n-first'} kernel={ctx.last_kernel} "
              f"err nopair {e0.max():.2e} pair {e1.max():.2e} bad {bad.sum()} of {bad.size}", flush=True)
        if bad.any():
            rows, cols = np.nonzero(bad)
            print("  bad row tiles(256):", np.unique(rows // 256)[:20], "count", len(np.unique(rows // 256)))
            print("  bad 128-halves:", np.unique((rows // 128) % 2), " bad col tiles(128):", np.unique(cols // 128), " col 64-halves:", np.unique((cols // 64) % 2))
            print("  bad rows mod 128 span:", rows.min() % 128, (rows % 128).max(), " cols mod 64:", (cols % 64).min(), (cols % 64).max())
            r0, c0 = rows[0], cols[0]
            print("  first bad", r0, c0, res[1][r0, c0], ref[r0, c0], " ratio", res[1][r0, c0] / ref[r0, c0])
            # is the pair result a sum over a subset of k chunks?
            ach = a.reshape(-1, K).astype(np.complex128) if False else None
