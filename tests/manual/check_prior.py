"""diagnostic: Rao-Teh sweeps WITHOUT observations must leave the prior invariant"""
import numpy as np, sys
sys.path.insert(0, '/root/repo')
import experiments
from smjp_b200.raoteh import raoteh_batch
from smjp_b200.chains import DeviceChains
from smjp_b200 import runtime
from smjp_b200.smjp_utils import sMJPDataWrapper, create_upperbound_scale
from oracle import smjp_oracle as O
m = experiments.build_experiment_2_model()
if len(sys.argv) > 1 and sys.argv[1] == 'shape1':
    s1 = np.ones((3, 3))
    m['hazard_A'].shape_mat = s1; m['hazard_B'].shape_mat = s1
    m['hazard_B'].scale_mat = create_upperbound_scale(s1, np.ones((3, 3)), 2)
nodata = sMJPDataWrapper(data=[], time=[])
C = 8000
eng = runtime.get_engine()
for mode in ('exact', 'reference'):
    out = raoteh_batch(C, 1, m['state_space'], 2.0, m['emission'], nodata, m['pi_0'], m['hazard_A'], m['hazard_B'],
                       seed=3, burn_in=40, candidate_mode=mode)
    print('raoteh-nodata', mode, 'time', out['time_in_state'][0].mean(0), 'jumps', out['jumps'][0].mean(0))
ch = DeviceChains(eng, C, np.zeros(0)).init_prior(np.ones(3) / 3, seed=77)
V, T = ch.paths_host()
tt = np.array([O.chain_metrics(v, t, 3)[0] for v, t in zip(V, T)]); jj = np.array([O.chain_metrics(v, t, 3)[1] for v, t in zip(V, T)])
print('prior          time', tt.mean(0), 'jumps', jj.mean(0), 'len', np.mean([len(v) for v in V]))
V2, T2 = out['paths']
print('raoteh len', np.mean([len(v) for v in V2]))
