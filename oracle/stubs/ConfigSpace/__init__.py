"""Stand-in for ConfigSpace==0.4.13 (TEST INFRASTRUCTURE ONLY): just enough for
/root/reference/agents/DDQN_vary.py:26-59 and DuelingDDQN_vary.py:24-75.  Samplers restate the published
semantics: log-uniform floats are exp(U(ln lo, ln hi)); log-uniform ints sample the float on
[lo-0.49999, hi+0.49999] in log space and round; plain ints are uniform on {lo..hi}.  The exact
ConfigSpace stream is not reproducible."""
import math

import numpy as np

from . import hyperparameters  # noqa: F401


class Configuration(dict):
    pass


class ConfigurationSpace(object):
    def __init__(self, seed=None):
        self._hps = []
        self.random = np.random.RandomState(seed)
        # test hook: uniform source u() -> [0,1)
        self.uniform_hook = None

    def add_hyperparameter(self, hp):
        self._hps.append(hp)
        return hp

    def sample_configuration(self):
        cfg = Configuration()
        for hp in self._hps:
            u = self.uniform_hook() if self.uniform_hook is not None else self.random.random_sample()
            cfg[hp.name] = hp._from_unit(u)
        return cfg
