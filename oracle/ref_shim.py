"""TEST INFRASTRUCTURE ONLY -- import shim for the *unmodified* reference (gauenk/smjp).

Used by ``oracle/gen_golden*.py`` (in the authoring container, where /root/reference exists) to produce the
committed fixtures under ``tests/golden/``, and by ``bench.py --impl reference`` to time the unmodified
reference beside the NumPy port: ``make -C oracle`` byte-compiles the reference's (pure-Python) modules from
where they lie under /root/reference into the git-ignored ``oracle/_ref/pkg/*.pycbin`` (bytecode only, no source is
copied), which travels to the GPU box with the snapshot and is imported sourceless.  Nothing in the product
package ``smjp_b200/`` imports this module.

What the shim does, without touching a single reference source line:
  * restores the numpy-1 aliases the reference uses (np.int, np.float, np.infty, np.math);
  * registers a minimal ``easydict.EasyDict`` and empty matplotlib/seaborn/imageio stubs;
  * puts the reference root on sys.path and silences its prints;
  * offers ``UniformSeam``: a stand-in for the module-global ``npr`` inside ``pkg.fast_hmm``
    whose ``multinomial(1, p)`` does inverse-CDF on the next host-supplied uniform, so that
    ``backward_sampling_hmm`` (fast_hmm.py:104-191) becomes a deterministic function of the
    uniforms.  Draw order in the reference is last grid step first (fast_hmm.py:125, :181).
"""
import builtins
import contextlib
import io
import math
import os
import sys
import types

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("SMJP_REFERENCE_ROOT") or (
    "/root/reference" if os.path.isfile("/root/reference/pkg/fast_hmm.py") else os.path.join(_HERE, "_ref"))


BYTECODE_EXT = ".pycbin"  # `make -C oracle`: py_compile output under a name snapshot tools do not filter out (*.pyc is)


def reference_available():
    """the sources (/root/reference) or their bytecode (oracle/_ref/pkg/*.pycbin, `make -C oracle`; imported sourceless)"""
    return any(os.path.isfile(os.path.join(REFERENCE_ROOT, "pkg", "fast_hmm" + ext)) for ext in (".py", BYTECODE_EXT))


class _BytecodeFinder:
    """sys.meta_path finder for the byte-compiled reference: package `pkg` and its modules from <root>/pkg/*.pycbin"""

    def __init__(self, root):
        self.dir = os.path.join(root, "pkg")

    def find_spec(self, fullname, path=None, target=None):
        import importlib.machinery
        import importlib.util
        if fullname == "pkg":
            f = os.path.join(self.dir, "__init__" + BYTECODE_EXT)
            if os.path.isfile(f):
                return importlib.util.spec_from_loader("pkg", importlib.machinery.SourcelessFileLoader("pkg", f),
                                                       origin=f, is_package=True)
        elif fullname.startswith("pkg.") and fullname.count(".") == 1:
            f = os.path.join(self.dir, fullname[4:] + BYTECODE_EXT)
            if os.path.isfile(f):
                return importlib.util.spec_from_loader(fullname, importlib.machinery.SourcelessFileLoader(fullname, f), origin=f)
        return None


class _EasyDict(dict):
    """Attribute-access dict, enough for distributions.py:1 ``from easydict import EasyDict``."""

    def __init__(self, d=None, **kw):
        super().__init__()
        if d is None:
            d = {}
        for k, v in dict(d, **kw).items():
            self[k] = v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def _install_stubs():
    if not hasattr(np, "int"):
        np.int = int
    if not hasattr(np, "float"):
        np.float = float
    if not hasattr(np, "infty"):
        np.infty = np.inf
    if not hasattr(np, "math"):
        np.math = math
    if "easydict" not in sys.modules:
        m = types.ModuleType("easydict")
        m.EasyDict = _EasyDict
        sys.modules["easydict"] = m
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.cm", "seaborn", "imageio"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                stub = types.ModuleType(name)
                stub.use = lambda *a, **k: None
                sys.modules[name] = stub
    if os.path.isfile(os.path.join(REFERENCE_ROOT, "pkg", "fast_hmm.py")):
        if REFERENCE_ROOT not in sys.path:
            sys.path.insert(0, REFERENCE_ROOT)
    elif not any(isinstance(f, _BytecodeFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _BytecodeFinder(REFERENCE_ROOT))


@contextlib.contextmanager
def quiet():
    """The reference prints inside its hot loop (fast_hmm.py:237); swallow it."""
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        yield


def load_reference():
    """Import the unmodified reference modules; returns a namespace of the ones on the path."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    _install_stubs()
    with quiet():
        import pkg.utils as ref_utils
        import pkg.fast_hmm as ref_fast_hmm
        import pkg.hidden_markov_model as ref_hmm
        import pkg.distributions as ref_dist
        import pkg.smjp_utils as ref_smjp
        import pkg.mcmc_utils as ref_mcmc
        import pkg.pmcmc as ref_pmcmc
        import pkg.raoteh as ref_raoteh
    return types.SimpleNamespace(utils=ref_utils, fast_hmm=ref_fast_hmm, hmm=ref_hmm,
                                 dist=ref_dist, smjp=ref_smjp, mcmc=ref_mcmc,
                                 pmcmc=ref_pmcmc, raoteh=ref_raoteh)


class UniformSeam:
    """Drop-in for ``numpy.random`` inside pkg.fast_hmm: categorical draws by inverse CDF.

    ``multinomial(1, p)`` returns the one-hot vector of ``searchsorted(cumsum(p), u, 'right')``
    for the next uniform ``u`` of the supplied sequence (clamped to the last positive entry).
    """

    def __init__(self, uniforms):
        self.uniforms = list(uniforms)
        self.used = 0

    def multinomial(self, n, p, size=None):
        assert n == 1 and size is None
        u = self.uniforms[self.used]
        self.used += 1
        p = np.asarray(p, dtype=np.float64)
        cdf = np.cumsum(p)
        idx = int(np.searchsorted(cdf, u, side="right"))
        last_pos = int(np.nonzero(p > 0)[0][-1])
        idx = min(idx, last_pos)
        out = np.zeros(len(p), dtype=np.int64)
        out[idx] = 1
        return out


def build_reference_model(ref, state_space, shape_mat, scale_mat, omega, t_end, obs_times, obs_data,
                          likelihood_power=1.0):
    """Model construction exactly as experiments.py:172-222 does it, with the reference classes."""
    S = ref.smjp
    shape_mat = np.asarray(shape_mat, dtype=np.float64)
    scale_mat = np.asarray(scale_mat, dtype=np.float64)
    scale_tilde = S.create_upperbound_scale(shape_mat, scale_mat, omega)
    scale_hat = S.create_upperbound_scale(shape_mat, scale_mat, omega - 1)
    hazard_A = S.smjpHazardFunction(state_space, shape_mat, scale_mat)
    hazard_B = S.smjpHazardFunction(state_space, shape_mat, scale_tilde, omega=omega)
    hazard_A_hat = S.smjpHazardFunction(state_space, shape_mat, scale_hat)
    pp_A_hat = S.PoissonProcess(state_space, None, hazard_A_hat, {"shape": shape_mat, "scale": scale_hat})
    pp_B = S.PoissonProcess(state_space, None, hazard_B, {"shape": shape_mat, "scale": scale_tilde})
    emission = S.smjpEmission(state_space, pp_B, t_end, likelihood_power)
    data = S.sMJPDataWrapper(data=np.asarray(obs_data), time=np.asarray(obs_times, dtype=np.float64))
    pi_0 = ref.dist.MultinomialDistribution({"prob_vector": np.ones(len(state_space)) / len(state_space),
                                             "translation": state_space})
    return types.SimpleNamespace(hazard_A=hazard_A, hazard_B=hazard_B, hazard_A_hat=hazard_A_hat,
                                 pp_A_hat=pp_A_hat, pp_B=pp_B, emission=emission, data=data, pi_0=pi_0,
                                 state_space=state_space, t_end=t_end, omega=omega)


def reference_ffbs(ref, model, W, uniforms):
    """Run the unmodified smjp_ffbs (smjp_utils.py:261-330) with the RNG seam installed.

    ``uniforms[i]`` is consumed by the draw at grid step i (the reference draws i = n-1 first).
    Returns (log_alphas [n, K*n], samples [n] int, V float64[m], T float64[m]).
    """
    W = [float(w) for w in W]
    n = len(W) - 1
    seam = UniformSeam(list(uniforms)[::-1][:n] if len(uniforms) == n else list(uniforms)[::-1])
    saved = ref.fast_hmm.npr
    captured = {}
    orig_bs = ref.fast_hmm.backward_sampling_hmm

    def capture_bs(*a, **kw):
        captured["alphas"] = np.array(kw["alphas"], dtype=np.float64)
        s, t = orig_bs(*a, **kw)
        captured["samples"] = np.array(s[0], dtype=np.int64)
        return s, t

    ref.fast_hmm.npr = seam
    ref.fast_hmm.backward_sampling_hmm = capture_bs
    try:
        with quiet():
            V, T, prob = ref.smjp.smjp_ffbs(W, model.data, model.state_space, model.hazard_A,
                                            model.hazard_B, model.emission, model.t_end, "oracle")
    finally:
        ref.fast_hmm.npr = saved
        ref.fast_hmm.backward_sampling_hmm = orig_bs
    assert seam.used == n, (seam.used, n)
    return captured["alphas"], captured["samples"], np.asarray(V, dtype=np.float64), \
        np.asarray(T, dtype=np.float64), float(prob)
