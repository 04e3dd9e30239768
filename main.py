"""Entry point mirroring the reference's main.py: exact MCMC for a semi-Markov jump process with
Weibull hazards (Rao & Teh), on the B200 engine.  `python main.py [number_of_samples]`."""
import sys

import numpy as np

from experiments import experiment_2

if __name__ == "__main__":
    np.set_printoptions(precision=3)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 600
    inference = ["trajectory", "parameters"]  # main.py:21
    experiment_2(likelihood_power=1.0, inference=inference, number_of_samples=n)
