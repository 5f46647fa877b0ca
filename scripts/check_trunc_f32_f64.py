"""Exhaustive check behind DESIGN.md's decoder statement: for EVERY float32 p in [0, 1], does
trunc(float32(p * scale)) (the encoder's residual, compress.py:307,311) equal trunc(float64(p) * scale) (the reference
decoder's reconstruction, decompress.py:252,269)?

    python scripts/check_trunc_f32_f64.py 255      ->  0 mismatches   (the 8-bit container: both forms are one function)
    python scripts/check_trunc_f32_f64.py 65535    ->  32895 mismatches (container v2 must state the decoder with the
                                                       float32 product, as the encoder does)
Takes ~1 minute per scale (1.07e9 values)."""
import numpy as np, sys
scale = float(sys.argv[1])
bad = 0; first = []
N = 0x3F800000 + 1
step = 1 << 24
for a in range(0, N, step):
    b = min(N, a + step)
    bits = np.arange(a, b, dtype=np.uint32)
    p = bits.view(np.float32)
    t32 = (p * np.float32(scale)).astype(np.int64)
    t64 = (p.astype(np.float64) * scale).astype(np.int64)
    d = np.nonzero(t32 != t64)[0]
    bad += d.size
    if d.size and len(first) < 5:
        first += [(float(p[i]), int(t32[i]), int(t64[i])) for i in d[:5]]
print(scale, "mismatches", bad, first[:5])
