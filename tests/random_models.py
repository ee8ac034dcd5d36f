"""Randomly generated compartmental models (seeded), written against the BaseModel API only, so
that the same class body runs on the reference's basepop and on popdynamics_b200.basepop.  They
exercise the flow compiler beyond the six example models: every flow kind, dedup of repeated
(from, to) pairs, vars that read other vars, divisions, many rows (more than 32 tau-leap rows)."""
import numpy as np


def random_model_class(basepop_module, old_div):
    class RandomModel(basepop_module.BaseModel):
        def __init__(self, seed, n_comp=4, scale=200, n_var=3, n_fixed=4, n_entry=1,
                     n_infdeath=1, background=True):
            basepop_module.BaseModel.__init__(self)
            rng = np.random.default_rng(seed)
            self.names = ['c%d' % i for i in range(n_comp)]
            for name in self.names:
                self.set_compartment(name, int(rng.integers(0, scale)))
            self.set_compartment(self.names[0], int(scale))
            self.var_defs = []
            for k in range(n_var):
                a, b = rng.integers(0, n_comp, 2)
                self.set_param('beta%d' % k, float(rng.uniform(0.05, 0.6)))
                self.var_defs.append(('v%d' % k, 'beta%d' % k, self.names[a], bool(k % 2),
                                      self.names[a], self.names[b]))
            self.fixed_defs = []
            for k in range(n_fixed):
                a, b = rng.choice(n_comp, 2, replace=(n_comp < 2))
                self.set_param('rate%d' % k, float(rng.uniform(0.01, 0.3)))
                self.fixed_defs.append((self.names[a], self.names[b], 'rate%d' % k))
            self.entry_defs = [(self.names[int(rng.integers(0, n_comp))], 'birth%d' % k)
                               for k in range(n_entry)]
            for k in range(n_entry):
                self.set_param('birth%d' % k, float(rng.uniform(0.0, 0.02)))
            self.infdeath_defs = []
            for k in range(n_infdeath):
                self.set_param('kill%d' % k, float(rng.uniform(0.0, 0.05)))
                self.infdeath_defs.append((self.names[int(rng.integers(0, n_comp))], 'kill%d' % k))
            self.set_param('mu', float(rng.uniform(0.001, 0.02)))
            self.background = background

        def calculate_vars(self):
            self.vars['population'] = sum(self.compartments.values())
            for k, (entry, label) in enumerate(self.entry_defs):
                self.vars[label] = self.params[label] * self.vars['population']
            for name, beta, comp, per_capita, _, _ in self.var_defs:
                value = self.params[beta] * self.compartments[comp]
                if per_capita:
                    value = old_div(value, self.vars['population'] + 1.0)
                else:
                    value = value * 0.01
                self.vars[name] = value
            if self.var_defs:
                # a var that reads other vars
                self.vars['v_sum'] = (self.vars[self.var_defs[0][0]] +
                                      0.5 * self.vars[self.var_defs[-1][0]])

        def set_flows(self):
            for comp, label in self.entry_defs:
                self.set_var_entry_rate_flow(comp, label)
            for k, (name, _, _, _, a, b) in enumerate(self.var_defs):
                if a != b:
                    self.set_var_transfer_rate_flow(a, b, name if k else 'v_sum')
            for a, b, label in self.fixed_defs:
                if a != b:
                    self.set_fixed_transfer_rate_flow(a, b, label)
            if self.background:
                self.set_background_death_rate('mu')
            for comp, label in self.infdeath_defs:
                self.set_infection_death_rate_flow(comp, label)
            # every compartment must appear in a flow (basepop.py:654-656)
            for i, name in enumerate(self.names):
                if len(self.names) > 1 and 'rate0' in self.params:
                    self.set_fixed_transfer_rate_flow(name, self.names[(i + 1) % len(self.names)],
                                                      'rate0')
                else:
                    self.set_var_entry_rate_flow(name, self.entry_defs[0][1])

    return RandomModel


CASES = [
    # seed, kwargs
    (1, dict(n_comp=1, n_var=0, n_fixed=0, n_entry=1, n_infdeath=1)),
    (2, dict(n_comp=2, n_var=1, n_fixed=1)),
    (3, dict(n_comp=4)),
    (4, dict(n_comp=5, n_var=4, n_fixed=6, n_entry=2, n_infdeath=2, scale=3000)),
    (5, dict(n_comp=7, n_var=6, n_fixed=9, n_entry=3, n_infdeath=4, background=False)),
    (6, dict(n_comp=9, n_var=9, n_fixed=14, n_entry=3, n_infdeath=6, scale=60)),   # > 32 rows
]
