"""TEST INFRASTRUCTURE ONLY -- fixtures for grids with a point within numpy.isclose() of t_end, produced by the
UNMODIFIED reference (/root/reference behind oracle/ref_shim.py):

    python -m oracle.gen_golden_final_rows

The emission drops its Poisson-event factor B_v on every interval whose right end "is" t_end
(smjp_utils.py:544-547, np.isclose: |w - T| <= 1e-8 + 1e-5 T), while the transition keeps A_vv'/B_v
(smjp_utils.py:409-412).  A candidate time that close to t_end (0.5 % of the cfg2 chain-sweeps) therefore gives
TWO OR MORE such rows, and the jump term out of the first of them no longer cancels against the emission -- the
case the round-1 kernels got wrong in the last message row.  Separate generator (own seed) so that the fixtures of
oracle/gen_golden.py keep their bytes.
"""
import numpy as np

from . import gen_golden, ref_shim


def main():
    ref = ref_shim.load_reference()
    rng = np.random.default_rng(20261020)
    t_end = 2.0
    inner = np.sort(rng.uniform(0, 1.9, 7))
    # one grid point within isclose of t_end: rows n-2 and n-1 are both "final"
    W1 = np.concatenate([[0.0], inner, [t_end - 1.0e-5, t_end]])
    gen_golden.ffbs_case(ref, "final_rows_2", [1, 2, 3], gen_golden.SHAPE_CFG1, np.ones((3, 3)), 2, t_end,
                         gen_golden.OBS_T_CFG1, gen_golden.OBS_Y_CFG1, 1.0, W1, rng.uniform(size=len(W1) - 1))
    # two such points, non-unit scales, Omega = 3: three final rows
    shape = rng.uniform(0.6, 3.0, (3, 3))
    scale = rng.uniform(0.5, 2.0, (3, 3))
    W2 = np.concatenate([[0.0], inner[:6], [t_end - 1.5e-5, t_end - 4.0e-6, t_end]])
    gen_golden.ffbs_case(ref, "final_rows_3", [1, 2, 3], shape, scale, 3, t_end,
                         gen_golden.OBS_T_CFG1, gen_golden.OBS_Y_CFG1, 1.0, W2, rng.uniform(size=len(W2) - 1))


if __name__ == "__main__":
    main()
