"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/shape_mle.npz by running the UNMODIFIED reference's
parameter-step helpers (pkg/raoteh.py:119-167: filter_T_by_state, esimtate_weibull_shape) on seeded prior
paths, in the authoring container:

    python -m oracle.gen_golden_params

Contents: the packed paths (V as 0-based indices, T, lens), and per chain and pooled over chains, for every
(v, v'): count of holding times, the reference's k-hat and its g(k-hat) (NaN where the list is empty -- the
reference then draws U(0.6, 3), raoteh.py:127-128)."""
import os

import numpy as np

from . import ref_shim
from . import smjp_oracle as O
from .gen_golden import OUT, SHAPE_CFG1


def main():
    ref = ref_shim.load_reference()
    K, t_end, C = 3, 40.0, 6
    labels = [1, 2, 3]
    paths = [O.prior_path(SHAPE_CFG1, np.ones((K, K)), np.ones(K) / K, t_end, seed=77, chain=c, honour_scale=True)
             for c in range(C)]
    khat = np.full((C + 1, K, K), np.nan)
    gval = np.full((C + 1, K, K), np.nan)
    count = np.zeros((C + 1, K, K), dtype=np.int64)
    pooled = {(a, b): [] for a in range(K) for b in range(K)}
    with ref_shim.quiet():
        for c, (V, T) in enumerate(paths):
            Vl = np.asarray([labels[v] for v in V], dtype=np.float64)  # the reference carries labels as floats
            for a in range(K):
                for b in range(K):
                    h = ref.raoteh.filter_T_by_state(np.asarray(T), Vl, labels[a], labels[b])
                    count[c, a, b] = len(h)
                    pooled[(a, b)] += list(h)
                    if len(h):
                        khat[c, a, b], gval[c, a, b] = ref.raoteh.esimtate_weibull_shape(h)
        for (a, b), h in pooled.items():
            count[C, a, b] = len(h)
            if len(h):
                khat[C, a, b], gval[C, a, b] = ref.raoteh.esimtate_weibull_shape(h)
    np.savez(os.path.join(OUT, "shape_mle.npz"), V=np.concatenate([p[0] for p in paths]).astype(np.int64),
             T=np.concatenate([p[1] for p in paths]), lens=np.array([len(p[0]) for p in paths], dtype=np.int64),
             K=np.int64(K), khat=khat, gval=gval, count=count, grid=np.float64(0.001), grid_max=np.float64(5.0))
    print("shape_mle: counts per chain", count[:C].sum(axis=(1, 2)), "pooled khat\n", khat[C])


if __name__ == "__main__":
    main()
