import numpy as np
# fit P(s) with atan(q) = q + q*s*P(s), s=q^2, q in [0,1]; minimise relative error of atan via Remez-like iterated reweighted LSQ
def fit(deg):
    q = np.cos(np.linspace(0, np.pi/2, 4001))[::-1]  # chebyshev-ish nodes on [0,1]
    q = q[q > 1e-4]
    s = q*q
    target = (np.arctan(q) - q) / (q*s)      # P(s)
    # error in atan relative: q*s*dP / atan(q) -> weight
    w = q*s/np.arctan(q)
    V = np.vander(s, deg+1, increasing=True)
    wt = np.ones_like(q)
    for it in range(60):
        A = V * (w*wt)[:, None]
        c, *_ = np.linalg.lstsq(A, target*w*wt, rcond=None)
        err = np.abs((V@c - target)*w)
        wt = wt * (1 + 2*err/err.max())  # Lawson
        wt /= wt.max()
    return c, err.max()
for deg in (6,7,8):
    c, e = fit(deg)
    print(deg, e, e/2**-24)
    c32 = c.astype(np.float32)
    # float32 evaluation test on dense grid
    q = np.linspace(0, 1, 2000001).astype(np.float32)
    s = (q*q).astype(np.float32)
    p = np.full_like(q, c32[-1])
    for k in range(deg-1, -1, -1):
        p = (p*s + c32[k]).astype(np.float32)   # not fused, slightly pessimistic
    t = (s*p).astype(np.float32)
    r = (q*t + q).astype(np.float32)
    ref = np.arctan(q.astype(np.float64))
    ulp = np.spacing(ref.astype(np.float32)).astype(np.float64)
    print('  max ulp err', np.max(np.abs(r.astype(np.float64)-ref)/ulp), [float(x) for x in c32])
