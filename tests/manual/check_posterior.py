import numpy as np, sys
sys.path.insert(0, '/root/repo')
import experiments
from smjp_b200.raoteh import raoteh_batch
from smjp_b200.engine import pf_host
from smjp_b200 import runtime
from smjp_b200.smjp_utils import emission_arrays, observation_arrays
from oracle import smjp_oracle as O
shape_override = None
if len(sys.argv) > 1 and sys.argv[1] == 'shape1':
    shape_override = np.ones((3, 3))
m = experiments.build_experiment_2_model()
if shape_override is not None:
    from smjp_b200.smjp_utils import create_upperbound_scale
    m['hazard_A'].shape_mat = shape_override
    m['hazard_B'].shape_mat = shape_override
    m['hazard_B'].scale_mat = create_upperbound_scale(shape_override, np.ones((3,3)), 2)
C = 4000
for mode in ('exact', 'reference'):
    out = raoteh_batch(C, 1, m['state_space'], 2.0, m['emission'], m['data'], m['pi_0'], m['hazard_A'], m['hazard_B'],
                       seed=3, burn_in=60, candidate_mode=mode, obs_combine='product')
    print('raoteh', mode, 'time', out['time_in_state'][0].mean(0), 'jumps', out['jumps'][0].mean(0))
eng = runtime.get_engine()
eng.set_model(m['hazard_A'].shape_mat, m['hazard_A'].scale_mat, 2.0, emission_arrays(m['emission']), 1.0, 2.0)
obs_t, obs_y = observation_arrays(m['data'], m['state_space'])
import sys as _s
for res in ('systematic','multinomial'):
  r = pf_host(eng, obs_t, obs_y, np.ones(3)/3, 8192, seed=9, C=C, resampler=res)
  tt = np.array([O.chain_metrics(V, T, 3)[0] for V, T in zip(r['V'], r['T'])])
  jj = np.array([O.chain_metrics(V, T, 3)[1] for V, T in zip(r['V'], r['T'])])
  print('pf N=8192', res, 'time', tt.mean(0), 'jumps', jj.mean(0), 'loglik mean', r['loglik'].mean())
